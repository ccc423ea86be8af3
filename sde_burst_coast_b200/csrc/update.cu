// update.cu - per-agent burst-coast update for swarms that do not fit one thread block:
// ParticleBurstCoast with the burst decision, wall avoidance, Euler step, boundary and burst
// re-arm (/root/reference/src/agents_dynamics.cpp:150-273).
//
// One multi-kernel step is TWO kernels that do not depend on each other:
//   * the rows kernel (pair_allpairs.cu / celllist.cu): the agents that start a burst this step - the only ones that
//     read the three-zone sums (IntCalcPrey's gate, agents_interact.cpp:178) - are swept against their candidates
//     and updated right there by sbc_update_agent (sbc_kernels.cuh);
//   * update_bulk_kernel (here): every other agent, one thread each, a plain stream over the 36-byte state
//     (pv 16 + hv 8 + force 8 + timers 4 read and written: 72 B per agent-step, HBM-bound).
// Both read start-of-step positions from S.pv and write the new ones to S.pv_nxt (Step()'s semantics,
// swarmdyn.cpp:143-204: interactions see the positions at the start of the step), both append the agents they
// re-arm to the row list of the NEXT step (three lists in rotation, so nobody has to wait for a list to be consumed):
// no compaction pass, no dependent launch between pair pass and update.
#include "sbc_kernels.cuh"

namespace {

template <int BCT>   // 0: periodic box, known at compile time (sbc_update_agent); SBC_BC_RUNTIME otherwise
__global__ void __launch_bounds__(256)
update_bulk_kernel(const __grid_constant__ DevParams P, DevState S, StepArgs A) {
    const uint32_t step = sbc_step_of(A);
    const int nxt = (int)((step + 1u) % 3u);
    const int lane = threadIdx.x & 31;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    sbc_trace(S, step, SBC_TR_BULK, true);
    // the list after next is idle during this step: clear its counter for the step that appends to it
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.active_count[(step + 2u) % 3u] = 0;
        S.active_count[4 + (step + 2u) % 3u] = 0;
    }
    bool mine = false, rearm = false;
    uint8_t al = 0;
    uint32_t tm = 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 hv = make_float2(0.f, 0.f), f = make_float2(0.f, 0.f);
    if (g < (size_t)A.n_update) {
        // every load is issued before the first one is looked at: one memory round trip per agent instead of three
        al = S.alive[g];
        tm = S.tm[g];
        p = S.pv[g];
        hv = S.hv[g];
        f = S.force[g];
        mine = al && sbc_tm_bin(tm, P) != P.burst_steps;   // rows that start a burst belong to the rows kernel
        rearm = mine && sbc_will_rearm(tm, P);
    }
    // warp-aggregated append to the rows of the next step, issued before the update: the counter's atomic returns
    // while the warp computes
    // the agent's cell key travels with the list entry (one round trip less for the row); while the cell order is being
    // rebuilt beside this kernel the key array is in flux, and the new key follows from the start-of-step position the
    // build is reading too (same arithmetic as cell_count_kernel)
    int key = 0;
    bool edge = false;
    if (rearm && S.cell_key) {
        if (A.fresh_cells && A.cell_ncx > 0) {   // periodic box: the grid is fixed, the host knows it
            const int cx = min(max((int)floorf(p.x * A.cell_inv_w), 0), A.cell_ncx - 1);
            const int cy = min(max((int)floorf(p.y * A.cell_inv_w), 0), A.cell_ncx - 1);
            key = cy * A.cell_ncx + cx;
        } else if (A.fresh_cells) {
            key = -1;   // grid fitted to the agents on the device: the row looks its key up
        } else {
            key = (int)S.cell_key[g];
        }
        edge = sbc_edge_cell(key, A);   // rows next to a slab edge are listed apart
    }
    const unsigned m = __ballot_sync(0xffffffffu, rearm && !edge);
    int base = 0, at_edge = -1;
    if (m && lane == 0) base = atomicAdd(S.active_count + nxt, __popc(m));
    if (rearm && edge) at_edge = atomicAdd(S.active_count + 4 + nxt, 1);
    if (mine) {
        RowAcc acc;
        row_acc_clear(acc);
        sbc_update_agent<BCT>(P, S, g, acc, tm, p, hv, f, step, A);
    }
    if (m) {
        base = __shfl_sync(0xffffffffu, base, 0);
        if (rearm && !edge) S.active_list[nxt][base + __popc(m & ((1u << lane) - 1u))] = make_int2((int)g, key);
    }
    if (at_edge >= 0) S.edge_list[nxt][at_edge] = make_int2((int)g, key);
    sbc_trace(S, step, SBC_TR_BULK, false);
}

// last node of a graph-replayed launch of `count` steps: the device-resident step index moves on
__global__ void step_advance_kernel(uint32_t *step_dev, uint32_t count) { *step_dev += count; }

// rows of a fresh state (after sbc_set_state, a full domain exchange, an ensemble launch ...): the agents with
// bin_step == burst_steps, four per thread, warp-aggregated append. The order of the list does not matter - every
// row writes only its own state.
__global__ void __launch_bounds__(256)
compact_active_kernel(const __grid_constant__ DevParams P, DevState S, int all_alive, int n_scan, int2 *list, int *count) {
    const size_t total = (size_t)n_scan;
    const size_t quads = (total + 3) / 4;
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // whole warps iterate together, so the scan below may use the full mask
    for (size_t qb = blockIdx.x * (size_t)blockDim.x + (threadIdx.x & ~31u); qb < quads; qb += stride) {
        const size_t q = qb + lane, g0 = 4 * q;
        unsigned act = 0;   // bit k: agent g0 + k starts a burst
        if (q < quads) {
            if (g0 + 3 < total) {
                const uint32_t al = *reinterpret_cast<const uint32_t *>(S.alive + g0);
                if (al && all_alive) {
                    for (int k = 0; k < 4; k++) act |= ((al >> (8 * k)) & 0xffu) ? (1u << k) : 0u;
                } else if (al) {
                    const uint4 t = *reinterpret_cast<const uint4 *>(S.tm + g0);
                    act |= ((al & 0xffu) && sbc_tm_bin(t.x, P) == P.burst_steps) ? 1u : 0u;
                    act |= ((al & 0xff00u) && sbc_tm_bin(t.y, P) == P.burst_steps) ? 2u : 0u;
                    act |= ((al & 0xff0000u) && sbc_tm_bin(t.z, P) == P.burst_steps) ? 4u : 0u;
                    act |= ((al & 0xff000000u) && sbc_tm_bin(t.w, P) == P.burst_steps) ? 8u : 0u;
                }
            } else {
                for (int k = 0; k < 4 && g0 + k < total; k++)
                    if (S.alive[g0 + k] && (all_alive || sbc_tm_bin(S.tm[g0 + k], P) == P.burst_steps)) act |= 1u << k;
            }
        }
        if (__ballot_sync(0xffffffffu, act != 0) == 0) continue;
        const int n_mine = __popc(act);
        int incl = n_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int base = 0;
        if (lane == 31) base = atomicAdd(count, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - n_mine;
        for (int k = 0; k < 4; k++)
            if ((act >> k) & 1u) list[base++] = make_int2((int)(g0 + k), S.cell_key ? (int)S.cell_key[g0 + k] : 0);
    }
}

__global__ void clear_acc_kernel(DevState S) {
    const size_t total = (size_t)S.N * S.R;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total;
         g += (size_t)gridDim.x * blockDim.x) {
        float4 *a = reinterpret_cast<float4 *>(S.acc + g * 8);
        a[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        a[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<int4 *>(S.cnt + g * 4) = make_int4(0, 0, 0, 0);
    }
}

// the ensemble kernels' heading vector from phi (after steps on the multi-kernel path, which does not maintain it)
__global__ void refresh_cs_kernel(DevState S) {
    const size_t total = (size_t)S.N * S.R;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    float s, c;
    sincosf(S.hv[g].x, &s, &c);
    S.cs[g] = make_float2(c, s);
}

}  // namespace

int sbc_update_bulk_blocks(int n_update) { return (n_update + 255) / 256; }

void sbc_launch_update_bulk(const DevParams &P, const DevState &S, const StepArgs &A, cudaStream_t st) {
    const unsigned blocks = (unsigned)sbc_update_bulk_blocks(A.n_update);
    if (P.BC == 0) update_bulk_kernel<0><<<blocks, 256, 0, st>>>(P, S, A);
    else update_bulk_kernel<SBC_BC_RUNTIME><<<blocks, 256, 0, st>>>(P, S, A);
}

// compact the rows that start a burst (or, all_alive, every alive agent) of slots [0, n_scan) into list / count
// (count must be zero); `clear` also zeroes every row's accumulators (the test hook reads rows that were not active)
void sbc_launch_compact_rows(const DevParams &P, const DevState &S, bool clear, bool all_alive, int n_scan, int2 *list, int *count,
                             cudaStream_t st) {
    const size_t total = (size_t)S.N * S.R;
    const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
    if (clear) clear_acc_kernel<<<blocks, 256, 0, st>>>(S);
    compact_active_kernel<<<blocks, 256, 0, st>>>(P, S, all_alive ? 1 : 0, n_scan, list, count);
}

void sbc_launch_step_advance(uint32_t *step_dev, int count, cudaStream_t st) {
    step_advance_kernel<<<1, 1, 0, st>>>(step_dev, (uint32_t)count);
}

void sbc_launch_refresh_cs(const DevState &S, cudaStream_t st) {
    const size_t total = (size_t)S.N * S.R;
    refresh_cs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S);
}
