// pair_allpairs.cu - three-zone social interaction over ALL ordered pairs (metric neighbour rule),
// the GPU counterpart of InteractionGlobal -> IntCalcPrey -> SFM_4Zone
// (/root/reference/src/agents_interact.cpp:144-157,172-250; social_forces.cpp:25-52).
//
// This file holds the burst-gated pass: only rows with bin_step == burst_steps are active (a
// fraction burst_rate*dt of the agents per step). Rows are compacted first; one warp per
// active row, lanes stride over j with coalesced 16-byte loads, counters by warp reduction. The
// dense pass (every row active) lives in pair_dense.cu. Both write the same accumulator layout:
// acc[g][8] = rep.xy alg.xy att.xy (flee.xy untouched), cnt[g][4] = c_rep c_alg c_att.
#include "sbc_kernels.cuh"

namespace {

// ---- burst-gated rows ------------------------------------------------------------------------
// four consecutive agents per thread (one 4-byte load of the alive flags, two 16-byte loads of the timers):
// the pass is a plain stream over 9 bytes per agent. The order of the list does not matter - every row
// writes only its own accumulators.
__global__ void __launch_bounds__(256)
compact_active_kernel(DevState S, uint32_t burst_steps, int all_alive, int *list, int *count, uint32_t *step_inc) {
    // first kernel of a graph-replayed step: advance the device-resident step index for the kernels behind it
    if (step_inc && blockIdx.x == 0 && threadIdx.x == 0) *step_inc += 1u;
    const size_t total = (size_t)S.N * S.R;
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
                    const uint4 t01 = *reinterpret_cast<const uint4 *>(S.timers + g0);
                    const uint4 t23 = *reinterpret_cast<const uint4 *>(S.timers + g0 + 2);
                    act |= ((al & 0xffu) && t01.x == burst_steps) ? 1u : 0u;
                    act |= ((al & 0xff00u) && t01.z == burst_steps) ? 2u : 0u;
                    act |= ((al & 0xff0000u) && t23.x == burst_steps) ? 4u : 0u;
                    act |= ((al & 0xff000000u) && t23.z == burst_steps) ? 8u : 0u;
                }
            } else {
                for (int k = 0; k < 4 && g0 + k < total; k++)
                    if (S.alive[g0 + k] && (all_alive || S.timers[g0 + k].x == burst_steps)) act |= 1u << k;
            }
        }
        if (__ballot_sync(0xffffffffu, act != 0) == 0) continue;
        // warp-aggregated append: one atomic per warp iteration
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
            if ((act >> k) & 1u) list[base++] = (int)(g0 + k);
    }
}

constexpr int ROWS_THREADS = 128;  // one block per active row

__global__ void __launch_bounds__(ROWS_THREADS)
pair_rows_kernel(const __grid_constant__ DevParams P, DevState S, const int *__restrict__ list,
                 const int *__restrict__ count, int detect, uint32_t step_arg, const uint32_t *__restrict__ step_dev) {
    const uint32_t step = step_dev ? *step_dev : step_arg;
    __shared__ float shf[ROWS_THREADS / 32][6];
    __shared__ int shi[ROWS_THREADS / 32][3];
    const int lane = threadIdx.x & 31;
    const int n_rows = *count;
    const int N = S.N;
    unsigned n_amb = 0;
    for (int k = blockIdx.x; k < n_rows; k += gridDim.x) {
        const int g = list[k];
        const int rep = g / N, i = g - rep * N;
        const float4 *pv = S.pv + (size_t)rep * N;
        const float4 pi = pv[i];
        RowAcc A;
        row_acc_clear(A);
        for (int j = threadIdx.x; j < N; j += ROWS_THREADS) {
            if (j == i) continue;
            const float4 pj = pv[j];
            float dx = pj.x - pi.x, dy = pj.y - pi.y;
            if (P.BC == 0) {
                dx = sbc_delta_pbc(pi.x, pj.x, P);
                dy = sbc_delta_pbc(pi.y, pj.y, P);
            }
            const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
            bool amb;
            int z = sbc_zone_fast(d2, P, amb);
            if (amb) {
                z = sbc_zone_exact(pi.x, pi.y, pj.x, pj.y, P);
                n_amb++;
            }
            if (P.voronoi && z < 3) {
                DelaunayEdge e;
                sbc_delaunay_begin(e, pi.x, pi.y, pj.x, pj.y);
                for (int kk = 0; kk < N; kk++) {
                    const float4 pk = pv[kk];
                    if (kk != i && kk != j && pk.x == pk.x) sbc_delaunay_add(e, pk.x, pk.y);
                }
                if (!sbc_delaunay_end(e)) z = 3;
            }
            const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
            if (z == 0) {
                A.rep_x = __fmaf_rn(-dx, rinv, A.rep_x);
                A.rep_y = __fmaf_rn(-dy, rinv, A.rep_y);
                A.c_rep++;
            } else if (z == 1) {
                A.alg_x += pj.z;
                A.alg_y += pj.w;
                A.c_alg++;
            } else if (z == 2) {
                A.att_x = __fmaf_rn(dx, rinv, A.att_x);
                A.att_y = __fmaf_rn(dy, rinv, A.att_y);
                A.c_att++;
            }
        }
        sbc_block_reduce_row<ROWS_THREADS / 32>(A, shf, shi);
        if (threadIdx.x == 0) {
            float *acc = S.acc + (size_t)g * 8;
            int *cnt = S.cnt + (size_t)g * 4;
            acc[0] = A.rep_x; acc[1] = A.rep_y; acc[2] = A.alg_x; acc[3] = A.alg_y;
            acc[4] = A.att_x; acc[5] = A.att_y;
            cnt[0] = A.c_rep; cnt[1] = A.c_alg; cnt[2] = A.c_att;
        }
        if (detect && threadIdx.x < 32) sbc_detect_row_warp(P, S, (size_t)g, pi, step, lane);
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(S.stats + 3, (unsigned long long)n_rows * (unsigned long long)(N - 1));
}

// test hook: neighbour ids of every row in ascending j (same zone functions as the kernels above)
__global__ void nn_fill_kernel(const __grid_constant__ DevParams P, DevState S, int all_rows,
                               const long long *__restrict__ nn_ptr, int *__restrict__ nn_idx) {
    const size_t total = (size_t)S.N * S.R;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int N = S.N;
    const int rep = (int)(g / N), i = (int)(g - (size_t)rep * N);
    if (!S.alive[g]) return;
    if (!all_rows && S.timers[g].x != P.burst_steps) return;
    const float4 *pv = S.pv + (size_t)rep * N;
    const float4 pi = pv[i];
    long long w = nn_ptr[g];
    for (int j = 0; j < N; j++) {
        if (j == i) continue;
        const float4 pj = pv[j];
        float dx = pj.x - pi.x, dy = pj.y - pi.y;
        if (P.BC == 0) {
            dx = sbc_delta_pbc(pi.x, pj.x, P);
            dy = sbc_delta_pbc(pi.y, pj.y, P);
        }
        const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
        bool amb;
        int z = sbc_zone_fast(d2, P, amb);
        if (amb) z = sbc_zone_exact(pi.x, pi.y, pj.x, pj.y, P);
        if (P.voronoi && z < 3) {
            DelaunayEdge e;
            sbc_delaunay_begin(e, pi.x, pi.y, pj.x, pj.y);
            for (int k = 0; k < N; k++) {
                const float4 pk = pv[k];
                if (k != i && k != j && pk.x == pk.x) sbc_delaunay_add(e, pk.x, pk.y);
            }
            if (!sbc_delaunay_end(e)) z = 3;
        }
        if (z < 3) nn_idx[w++] = j;
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

}  // namespace

// compact the rows that start a burst this step into S.active_list / S.active_count; `clear` also
// zeroes every row's accumulators (only the test hook reads rows that were not active)
void sbc_launch_compact_rows(const DevParams &P, const DevState &S, bool clear, cudaStream_t st, bool all_alive,
                             bool count_is_zero, uint32_t *step_inc) {
    const size_t total = (size_t)S.N * S.R;
    const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
    if (!count_is_zero) cudaMemsetAsync(S.active_count, 0, sizeof(int), st);
    if (clear) clear_acc_kernel<<<blocks, 256, 0, st>>>(S);
    compact_active_kernel<<<blocks, 256, 0, st>>>(S, P.burst_steps, all_alive ? 1 : 0, S.active_list, S.active_count, step_inc);
}

void sbc_launch_pair_rows(const DevParams &P, const DevState &S, cudaStream_t st, int detect, long long step,
                          const uint32_t *step_dev) {
    // grid: enough blocks for the synchronised first burst (every row active), few enough to be cheap
    // to launch when only a handful of rows start a burst
    const size_t total = (size_t)S.N * S.R;
    const int blocks = (int)min((size_t)148 * 8, max((size_t)16, total / 64));
    pair_rows_kernel<<<blocks, ROWS_THREADS, 0, st>>>(P, S, S.active_list, S.active_count, detect, (uint32_t)step, step_dev);
}

void sbc_launch_nn_fill(const DevParams &P, const DevState &S, int all_rows, const long long *nn_ptr,
                        int *nn_idx, cudaStream_t st) {
    const size_t total = (size_t)S.N * S.R;
    nn_fill_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(P, S, all_rows, nn_ptr, nn_idx);
}
