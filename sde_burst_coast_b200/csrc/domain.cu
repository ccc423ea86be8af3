// domain.cu - spatial domain decomposition of ONE large periodic swarm over several handles (one per
// GPU): slabs along x, halo of the agents within att_range of a slab edge, migration of agents that
// crossed an edge (BASELINE config C5). The reference has no counterpart (single
// process); what must be preserved is Step()'s semantics (/root/reference/src/swarmdyn.cpp:143-204):
// interactions see every agent's position at the start of the step, every agent is updated once,
// decisions depend on (seed, GLOBAL agent id, step) only.
//
// Slots of a handle: [0, own_cap) owned agents (alive) or free slots (alive = 0, x = NaN);
// [own_cap, N) ghosts = copies of neighbours' edge agents (alive = 0 but x valid: visible as
// interaction partners j, never rows, never updated).
//
// ONE exchange per step, after the update:
//   pack   : every owned agent is classified in one ordered pass - it LEFT the slab through the left /
//            right edge (full-state record, slot freed) or it stays and lies within the halo band of
//            an edge (position record). Both kinds go into one fixed-size buffer per side
//            [header | migrants | halo], so no host round trip sizes a message.
//   unpack : incoming migrants take free slots (a device-resident stack, lowest slot first), incoming
//            halo records become ghosts, and so do the agents this rank just sent away - they sit just
//            across the edge, exactly where the neighbour's halo band would report them.
// REFRESH exchanges. The cell order of a handle serves several steps (Verlet skin, celllist.cu), and so
// does the slot layout: between two cell builds no agent changes rank and no ghost changes slot. The
// band mirrored on the neighbour is att_range + skin wide, the full exchange records which slots it sent
// to each side (plus the migrants it received from that side - the sender keeps those as ghosts), and
// the exchanges in between only carry the current {x, y, vx, vy} of exactly those slots, in the same
// order, into the same ghost slots. An agent that crosses an edge meanwhile stays with its old owner
// (at most skin/2 outside the slab) until the next full exchange.
// Packing order is by slot index (block counts -> scan -> ordered write) and arrival order is fixed
// (left buffer, right buffer), so results do not depend on thread scheduling. The transport itself
// (NCCL send/recv over NVLink, or a loop-back copy in tests) belongs to the host side
// (sde_burst_coast_b200/domain.py).
#include "sbc_kernels.cuh"

namespace {

struct __align__(16) DomRecord {  // 64 bytes
    float4 pv;
    float2 hv;       // phi, vproj
    float2 force;
    uint32_t tm;     // packed burst timers
    uint32_t spare[3];
    uint32_t gid;
    uint32_t pad[3];
};
static_assert(sizeof(DomRecord) == 64, "DomRecord must be 64 bytes");

constexpr int CLS_THREADS = 256;
constexpr int CLS_PER_BLOCK = 1024;  // slots per block
enum { CAT_LEAVE_L = 0, CAT_LEAVE_R = 1, CAT_HALO_L = 2, CAT_HALO_R = 3, CAT_FREE = 4, NCAT = 5 };

// category mask of one owned slot: bit c set = the slot belongs to category c
__device__ __forceinline__ unsigned classify(const DevState &S, int g, float x_lo, float x_hi, float L, float halo) {
    if (!S.alive[g]) return 0u;
    const float x = S.pv[g].x;
    if (x >= x_lo && x < x_hi) {
        unsigned m = 0;
        if (x < x_lo + halo) m |= 1u << CAT_HALO_L;
        if (x >= x_hi - halo) m |= 1u << CAT_HALO_R;
        return m;
    }
    // outside the slab: the neighbour it is closer to along the ring gets it
    float dl = x_lo - x; if (dl < 0.f) dl += L;
    float dr = x - x_hi; if (dr < 0.f) dr += L;
    return (dl <= dr) ? (1u << CAT_LEAVE_L) : (1u << CAT_LEAVE_R);
}

__global__ void __launch_bounds__(CLS_THREADS)
dom_count_kernel(DevState S, int own_cap, float x_lo, float x_hi, float L, float halo, int *__restrict__ block_counts) {
    __shared__ int sh[4];
    if (threadIdx.x < 4) sh[threadIdx.x] = 0;
    __syncthreads();
    int c[4] = {0, 0, 0, 0};
    const int base = blockIdx.x * CLS_PER_BLOCK;
    for (int k = threadIdx.x; k < CLS_PER_BLOCK; k += CLS_THREADS) {
        const int g = base + k;
        if (g >= own_cap) break;
        const unsigned m = classify(S, g, x_lo, x_hi, L, halo);
#pragma unroll
        for (int q = 0; q < 4; q++) c[q] += (m >> q) & 1u;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int v = __reduce_add_sync(0xffffffffu, c[q]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[q], v);
    }
    __syncthreads();
    if (threadIdx.x < 4) block_counts[blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x];
}

// exclusive scan of the per-block counts (one block), headers of the two send buffers, free-stack base
__global__ void __launch_bounds__(1024)
dom_scan_kernel(int nblk, const int *__restrict__ block_counts, int *__restrict__ block_off, int *__restrict__ ctl,
                DomRecord *__restrict__ send_l, DomRecord *__restrict__ send_r, int max_mig, int max_halo) {
    __shared__ int warp_tot[32][4];
    __shared__ int carry[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        int v[4], inc[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            v[q] = (b < nblk) ? block_counts[b * 4 + q] : 0;
            inc[q] = v[q];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc[q], o);
                if (lane >= o) inc[q] += t;
            }
            if (lane == 31) warp_tot[warp][q] = inc[q];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int t = warp_tot[lane][q], s = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, s, o);
                    if (lane >= o) s += u;
                }
                warp_tot[lane][q] = s - t;  // exclusive prefix of the warp totals
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int excl = carry[q] + warp_tot[warp][q] + inc[q] - v[q];
            if (b < nblk) block_off[b * 4 + q] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
#pragma unroll
            for (int q = 0; q < 4; q++) carry[q] += warp_tot[31][q] + inc[q];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // ctl: [0] free_top, [1] free_base for this pack, [2..5] totals, [6] overflow events
        const int nl = carry[CAT_LEAVE_L], nr = carry[CAT_LEAVE_R], hl = carry[CAT_HALO_L], hr = carry[CAT_HALO_R];
        ctl[2] = nl; ctl[3] = nr; ctl[4] = hl; ctl[5] = hr;
        if (nl > max_mig || nr > max_mig || hl > max_halo || hr > max_halo) ctl[6] += 1;
        ctl[1] = ctl[0];
        ctl[0] += min(nl, max_mig) + min(nr, max_mig);  // leavers free their slots
        DomRecord h = {};
        h.gid = (uint32_t)min(nl, max_mig); h.pad[0] = (uint32_t)min(hl, max_halo);
        send_l[0] = h;
        h.gid = (uint32_t)min(nr, max_mig); h.pad[0] = (uint32_t)min(hr, max_halo);
        send_r[0] = h;
    }
}

// ordered write of the records: within a block by a ballot prefix, between blocks by the scanned offsets
__global__ void __launch_bounds__(CLS_THREADS)
dom_pack_kernel(DevState S, int own_cap, float x_lo, float x_hi, float L, float halo, const int *__restrict__ block_off,
                const int *__restrict__ ctl, int *__restrict__ free_stack, DomRecord *__restrict__ send_l,
                DomRecord *__restrict__ send_r, int max_mig, int max_halo, int *__restrict__ halo_list) {
    __shared__ int run[4];       // running offsets of this block, advanced chunk by chunk
    __shared__ int wcnt[CLS_THREADS / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) run[threadIdx.x] = block_off[blockIdx.x * 4 + threadIdx.x];
    __syncthreads();
    const int base = blockIdx.x * CLS_PER_BLOCK;
    const int free_base = ctl[1], n_leave_l = min(ctl[2], max_mig);
    const float qnan = __int_as_float(0x7fc00000);
    for (int k0 = 0; k0 < CLS_PER_BLOCK; k0 += CLS_THREADS) {
        const int g = base + k0 + threadIdx.x;
        const unsigned m = (g < own_cap) ? classify(S, g, x_lo, x_hi, L, halo) : 0u;
        int pos[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned bal = __ballot_sync(0xffffffffu, (m >> q) & 1u);
            pos[q] = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) wcnt[warp][q] = __popc(bal);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int off = run[q];
            for (int w = 0; w < warp; w++) off += wcnt[w][q];
            pos[q] += off;
        }
        if (m) {
            DomRecord r = {};
            r.pv = S.pv[g];
            r.gid = S.gid ? S.gid[g] : (uint32_t)g;
            if (m & ((1u << CAT_LEAVE_L) | (1u << CAT_LEAVE_R))) {
                const bool left = (m >> CAT_LEAVE_L) & 1u;
                const int p = left ? pos[CAT_LEAVE_L] : pos[CAT_LEAVE_R];
                if (p < max_mig) {
                    r.hv = S.hv[g];
                    r.force = S.force[g];
                    r.tm = S.tm[g];
                    (left ? send_l : send_r)[1 + p] = r;
                    // the agent now lives on the neighbour: free the slot (ordered push on the stack)
                    S.alive[g] = 0;
                    S.pv[g] = make_float4(qnan, qnan, 0.f, 0.f);
                    free_stack[free_base + (left ? p : n_leave_l + p)] = g;
                }
            } else {
                const int list_cap = max_halo + max_mig;
                if ((m >> CAT_HALO_L) & 1u) {
                    const int p = pos[CAT_HALO_L];
                    if (p < max_halo) { send_l[1 + max_mig + p] = r; halo_list[p] = g; }
                }
                if ((m >> CAT_HALO_R) & 1u) {
                    const int p = pos[CAT_HALO_R];
                    if (p < max_halo) { send_r[1 + max_mig + p] = r; halo_list[list_cap + p] = g; }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            int t = 0;
            for (int w = 0; w < CLS_THREADS / 32; w++) t += wcnt[w][threadIdx.x];
            run[threadIdx.x] += t;
        }
        __syncthreads();
    }
}

// incoming migrants -> free owned slots; ghosts = incoming halo (left, right) + my own outgoing migrants
__global__ void __launch_bounds__(256)
dom_unpack_kernel(DevState S, int own_cap, const DomRecord *__restrict__ recv_a, const DomRecord *__restrict__ recv_b,
                  const DomRecord *__restrict__ sent_l, const DomRecord *__restrict__ sent_r, int max_mig, int max_halo,
                  int *__restrict__ ctl, const int *__restrict__ free_stack, int *__restrict__ halo_list, int epoch) {
    if (ctl[21]) return;   // the wait for these records timed out: leave the slots alone
    const int ma = recv_a ? (int)recv_a[0].gid : 0, mb = recv_b ? (int)recv_b[0].gid : 0;
    const int ha = recv_a ? (int)recv_a[0].pad[0] : 0, hb = recv_b ? (int)recv_b[0].pad[0] : 0;
    const int sl = sent_l ? (int)sent_l[0].gid : 0, sr = sent_r ? (int)sent_r[0].gid : 0;
    const int free_top = ctl[0];
    const int ghost_cap = S.N - own_cap;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int own_l = min(ctl[4], max_halo), own_r = min(ctl[5], max_halo);   // my own halo lists
    if (t == 0) {
        if (ma + mb > free_top || ha + hb + sl + sr > ghost_cap) ctl[6] += 1;
        if ((recv_a && recv_a[0].pad[1] != 0u) || (recv_b && recv_b[0].pad[1] != 0u)) ctl[6] += 1;  // not a full record set
        ctl[7] = ma + mb;  // consumed by dom_commit_kernel
        ctl[10] = ha; ctl[11] = hb; ctl[12] = sl; ctl[13] = sr;
        ctl[14] = own_l + ma;   // the migrants from the left sit in my left band: the left neighbour mirrors them
        ctl[15] = own_r + mb;
    }
    if (t < ma + mb && t < free_top) {  // migrants
        const DomRecord r = (t < ma) ? recv_a[1 + t] : recv_b[1 + (t - ma)];
        const int g = free_stack[free_top - 1 - t];
        if (t < ma) halo_list[own_l + t] = g;
        else halo_list[(max_halo + max_mig) + own_r + (t - ma)] = g;
        S.pv[g] = r.pv;
        S.pv_dead[g] = r.pv;
        S.hv[g] = r.hv;
        S.force[g] = r.force;
        S.tm[g] = r.tm;
        S.alive[g] = 1;
        if (S.gid) S.gid[g] = r.gid;
    }
    if (t < ghost_cap) {  // ghosts
        const float qnan = __int_as_float(0x7fc00000);
        float4 pv = make_float4(qnan, qnan, 0.f, 0.f);
        uint32_t gid = 0xffffffffu;
        int k = t;
        if (k < ha) { pv = recv_a[1 + max_mig + k].pv; gid = recv_a[1 + max_mig + k].gid; }
        else if ((k -= ha) < hb) { pv = recv_b[1 + max_mig + k].pv; gid = recv_b[1 + max_mig + k].gid; }
        else if ((k -= hb) < sl) { pv = sent_l[1 + k].pv; gid = sent_l[1 + k].gid; }
        else if ((k -= sl) < sr) { pv = sent_r[1 + k].pv; gid = sent_r[1 + k].gid; }
        const int g = own_cap + t;
        S.pv[g] = pv;
        S.alive[g] = 0;
        if (S.gid) S.gid[g] = gid;
    }
}
__global__ void dom_commit_kernel(int *ctl, int epoch) {
    if (ctl[21]) return;
    ctl[0] = max(ctl[0] - ctl[7], 0);
    if (epoch >= 0) ctl[17] = epoch;   // the ghosts hold the records of this exchange
}

// refresh exchange: current {x, y, vx, vy} of the slots listed by the last full exchange, same order
__global__ void __launch_bounds__(256)
dom_refresh_pack_kernel(DevState S, const int *__restrict__ ctl, const int *__restrict__ halo_list, int list_cap,
                        DomRecord *__restrict__ send_l, DomRecord *__restrict__ send_r) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = min(ctl[14], list_cap), nr = min(ctl[15], list_cap);
    if (t == 0) {
        DomRecord h = {};
        h.pad[1] = 1u;   // marks a refresh buffer
        h.pad[0] = (uint32_t)nl; send_l[0] = h;
        h.pad[0] = (uint32_t)nr; send_r[0] = h;
    }
    if (t < nl) send_l[1 + t].pv = S.pv[halo_list[t]];
    if (t < nr) send_r[1 + t].pv = S.pv[halo_list[list_cap + t]];
}
// ghosts keep the slots the last full unpack gave them: [halo from a][halo from b][sent left][sent right];
// the agents I sent left now live on the left neighbour, which lists them behind its own right band
__global__ void __launch_bounds__(256)
dom_refresh_unpack_kernel(DevState S, int own_cap, int *__restrict__ ctl, const DomRecord *recv_a0, const DomRecord *recv_a1,
                          const DomRecord *recv_b0, const DomRecord *recv_b1, int epoch) {
    if (ctl[21]) return;
    const int par = epoch & 1;   // buffers by parity of the exchange count (the same pointer twice without an inbox)
    const DomRecord *recv_a = par ? recv_a1 : recv_a0, *recv_b = par ? recv_b1 : recv_b0;
    const int ha = ctl[10], hb = ctl[11], sl = ctl[12], sr = ctl[13];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        const bool ok_a = recv_a ? (recv_a[0].pad[1] == 1u && (int)recv_a[0].pad[0] == ha + sl) : (ha + sl == 0);
        const bool ok_b = recv_b ? (recv_b[0].pad[1] == 1u && (int)recv_b[0].pad[0] == hb + sr) : (hb + sr == 0);
        if (!ok_a || !ok_b) ctl[6] += 1;   // the ranks disagree about the kind or the content of this exchange
    }
    int k = t;
    const DomRecord *src = nullptr;
    if (k < ha) src = recv_a + 1 + k;
    else if ((k -= ha) < hb) src = recv_b + 1 + k;
    else if ((k -= hb) < sl) src = recv_a + 1 + ha + k;
    else if ((k -= sl) < sr) src = recv_b + 1 + hb + k;
    if (src && t < S.N - own_cap) S.pv[own_cap + t] = src->pv;
    if (t == 0) ctl[17] = epoch;   // (read by kernels launched behind this one)
}

// ---- exchange through peer memory (NVLink) ---------------------------------------------------------
// Every rank owns an INBOX: per side (from left / from right) two buffers (parity of the exchange epoch)
// and one arrival flag each. A rank pushes the used part of its send buffers straight into its
// neighbours' inboxes (16-byte stores over NVLink), then a one-thread kernel publishes the epoch in the
// neighbours' flags; the receiver's one-thread wait kernel holds its stream until both flags show the
// epoch. Producers never wait, so the ring cannot deadlock; two parities suffice because a rank can
// only push epoch e + 2 after it has consumed e + 1, which the neighbour sent after consuming e.
// Exchanges are numbered (one per step, full or refresh). The host passes the number to the kernels it launches
// eagerly; inside a replayed graph it is the device-resident step index plus a shift the host keeps in ctl[18].
// ctl[17] holds the number of the last exchange whose records are in the ghost slots.
__global__ void __launch_bounds__(256)
dom_push_kernel(const DomRecord *__restrict__ send_l, const DomRecord *__restrict__ send_r, DomRecord *dst_l0,
                DomRecord *dst_l1, DomRecord *dst_r0, DomRecord *dst_r1, int epoch, int max_mig,
                int max_halo) {
    const int par = epoch & 1;
    DomRecord *dst_l = par ? dst_l1 : dst_l0, *dst_r = par ? dst_r1 : dst_r0;
    const int per_side = (1 + max_mig + max_halo) * 4;   // 16-byte quarters
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 2 * per_side; q += gridDim.x * blockDim.x) {
        const int side = q >= per_side;
        const int k = q - side * per_side, rec = k >> 2;
        const DomRecord *src = side ? send_r : send_l;
        const DomRecord h = src[0];
        bool used;
        if (h.pad[1] == 1u) used = rec < 1 + (int)h.pad[0];                       // refresh: [header | n records]
        else used = rec < 1 + (int)h.gid || (rec >= 1 + max_mig && rec < 1 + max_mig + (int)h.pad[0]);
        if (used) reinterpret_cast<uint4 *>(side ? dst_r : dst_l)[k] = reinterpret_cast<const uint4 *>(src)[k];
    }
}
__global__ void dom_flag_kernel(volatile int *flag_l0, volatile int *flag_l1, volatile int *flag_r0, volatile int *flag_r1,
                                int epoch) {
    const int par = epoch & 1;
    __threadfence_system();
    *(par ? flag_l1 : flag_l0) = epoch;
    *(par ? flag_r1 : flag_r0) = epoch;
}
// Wall-clock nanoseconds (independent of the SM clock).
__device__ __forceinline__ unsigned long long dom_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Hold until both flags show `epoch`. A neighbour that does not answer within timeout_ns is reported, not waited
// for: ctl[21] (sticky) marks the handle's exchange as failed, every later unpack is skipped and the host
// side turns the mark into SBC_ERR_STATE (sbc_domain_step / sbc_domain_complete).
__device__ __forceinline__ bool dom_wait_flags(const volatile int *fa, const volatile int *fb, int epoch,
                                               unsigned long long timeout_ns) {
    // system-scope acquire loads: the flags are written by another GPU, and a plain volatile load may be answered by a
    // copy in this GPU's other L2 partition for several microseconds
    auto ld_sys = [](const volatile int *p) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    };
    const unsigned long long t0 = dom_now_ns();
    while (ld_sys(fa) < epoch || ld_sys(fb) < epoch) {
        if (dom_now_ns() - t0 > timeout_ns) return false;
        __nanosleep(50);
    }
    return true;
}
__global__ void dom_wait_kernel(const volatile int *flag_a0, const volatile int *flag_a1, const volatile int *flag_b0,
                                const volatile int *flag_b1, int *ctl, int epoch, unsigned long long timeout_ns) {
    const int par = epoch & 1;
    if (ctl[21]) return;   // an earlier exchange already failed
    if (!dom_wait_flags(par ? flag_a1 : flag_a0, par ? flag_b1 : flag_b0, epoch, timeout_ns)) {
        ctl[6] += 1000;
        ctl[21] = 1;
    }
}

// One refresh exchange in two kernels (the graph-replayed form): SEND gathers the listed positions straight into
// the neighbours' inboxes and the last block to finish publishes the exchange number; RECV holds every block until
// both neighbours' numbers arrived, scatters their positions into the ghost slots, and its last block announces in
// ctl[17] that the ghosts are in place - which is what the rows kernel of the running step waits for before it reads
// a candidate near a slab edge. `step_dev` / `step_off`: the exchange belongs to step (*step_dev + step_off) - 1, the
// one BEFORE the step in whose graph these kernels sit (nullptr: step_off is that number itself, eager launch).
__device__ __forceinline__ int dom_epoch_of(const int *ctl, const uint32_t *step_dev, uint32_t step_off) {
    const uint32_t step = step_dev ? *step_dev + step_off : step_off;
    return (int)step - 1 + ctl[18];
}
// The records of a refresh exchange are delivered by the threads that updated the agents (sbc_update_agent); the last
// node of the step's graph raises the neighbours' flags (dom_publish_kernel). The receive - a branch of the NEXT step's
// graph - holds until both neighbours' flags show exchange `epoch`, then scatters the records into the ghost slots.
__global__ void dom_publish_kernel(const int *ctl, const uint32_t *step_dev, uint32_t step_off, volatile int *flag_l0,
                                   volatile int *flag_l1, volatile int *flag_r0, volatile int *flag_r1) {
    const int epoch = dom_epoch_of(ctl, step_dev, step_off), par = epoch & 1;
    __threadfence_system();   // (the delivering kernels are complete: their stores are on their way ahead of these)
    *(par ? flag_l1 : flag_l0) = epoch;
    *(par ? flag_r1 : flag_r0) = epoch;
}
__device__ __forceinline__ void dom_recv_body(const DevState &S, int own_cap, int *ctl, int epoch, const DomRecord *recv_a0,
                                              const DomRecord *recv_a1, const DomRecord *recv_b0, const DomRecord *recv_b1,
                                              const volatile int *flag_a0, const volatile int *flag_a1,
                                              const volatile int *flag_b0, const volatile int *flag_b1,
                                              unsigned long long timeout_ns, unsigned *ticket, int bid, int nblk) {
    const int par = epoch & 1;
    __shared__ int arrived;
    sbc_trace(S, (uint32_t)epoch, SBC_TR_RECV, true);
    if (threadIdx.x == 0) {
        // ctl[21] of an earlier failure is read through volatile: block 0 of THIS launch may be setting it right now,
        // which is fine - every block runs its own wait and comes to the same verdict
        arrived = (*(volatile int *)(ctl + 21) == 0) &&
                  dom_wait_flags(par ? flag_a1 : flag_a0, par ? flag_b1 : flag_b0, epoch, timeout_ns);
        if (S.trace) {   // when the first / the last block of the receive saw both flags (kernel id SBC_TR_SEND)
            unsigned long long *w = S.trace + ((size_t)((uint32_t)epoch % SBC_TRACE_STEPS) * SBC_TR_KERNELS + SBC_TR_SEND) * 2;
            const unsigned long long now = sbc_now_ns();
            atomicMin(w, now);
            atomicMax(w + 1, now);
        }
        if (!arrived && bid == 0 && *(volatile int *)(ctl + 21) == 0) {
            ctl[6] += 1000;
            *(volatile int *)(ctl + 21) = 1;
        }
    }
    __syncthreads();
    if (arrived) {   // stale or partial records are never unpacked
        const DomRecord *recv_a = par ? recv_a1 : recv_a0, *recv_b = par ? recv_b1 : recv_b0;
        const int ha = ctl[10], hb = ctl[11], sl = ctl[12], sr = ctl[13];
        const int t = bid * blockDim.x + threadIdx.x;
        int k = t;
        const DomRecord *src = nullptr;
        if (k < ha) src = recv_a + 1 + k;
        else if ((k -= ha) < hb) src = recv_b + 1 + k;
        else if ((k -= hb) < sl) src = recv_a + 1 + ha + k;
        else if ((k -= sl) < sr) src = recv_b + 1 + hb + k;
        if (src && t < S.N - own_cap) {
            // the records were written by another GPU: read them past L1
            const float4 pv = __ldcg(reinterpret_cast<const float4 *>(&src->pv));
            S.pv[own_cap + t] = pv;
        }
    }
    sbc_trace(S, (uint32_t)epoch, SBC_TR_RECV, false);
    // the last block tells the rows kernel that runs beside this one that the ghosts are in place (also after a
    // time-out: the step must not hang; the failure is on record in ctl[21])
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(ticket, 1u);
        if (done == (unsigned)nblk - 1u) {
            *ticket = 0u;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ctl + 17), "r"(epoch) : "memory");
        }
    }
}
__global__ void __launch_bounds__(256)
dom_refresh_receive_kernel(DevState S, int own_cap, int *ctl, const uint32_t *step_dev, uint32_t step_off, const DomRecord *recv_a0,
                           const DomRecord *recv_a1, const DomRecord *recv_b0, const DomRecord *recv_b1,
                           const volatile int *flag_a0, const volatile int *flag_a1, const volatile int *flag_b0,
                           const volatile int *flag_b1, unsigned long long timeout_ns, volatile int *flag_l0,
                           volatile int *flag_l1, volatile int *flag_r0, volatile int *flag_r1) {
    const int epoch = dom_epoch_of(ctl, step_dev, step_off);
    // this rank's own records of the same exchange were delivered by the step that has just ended (sbc_update_agent):
    // raise the neighbours' flags first (flag_l0 == nullptr: a publish node behind that step has done it already)
    if (flag_l0 && blockIdx.x == 0 && threadIdx.x == 0) {
        const int par = epoch & 1;
        __threadfence_system();
        *(par ? flag_l1 : flag_l0) = epoch;
        *(par ? flag_r1 : flag_r0) = epoch;
    }
    dom_recv_body(S, own_cap, ctl, epoch, recv_a0, recv_a1, recv_b0, recv_b1, flag_a0, flag_a1, flag_b0, flag_b1, timeout_ns,
                  reinterpret_cast<unsigned *>(ctl + 22), blockIdx.x, gridDim.x);
}

// inverse of the refresh lists: where each listed slot's position goes in the records to the left / right neighbour
__global__ void dom_halo_idx_kernel(const int *__restrict__ ctl, const int *__restrict__ halo_list, int list_cap, int2 *halo_idx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = min(ctl[14], list_cap), nr = min(ctl[15], list_cap);
    if (t < nl) halo_idx[halo_list[t]].x = t;
    if (t < nr) halo_idx[halo_list[list_cap + t]].y = t;
}

// free stack of a fresh state: the free owned slots in descending order, so that pops take the lowest
// slot first. One block walks the slots from the top down, 1024 at a time, with an ordered compaction.
__global__ void __launch_bounds__(1024)
dom_init_stack_kernel(DevState S, int own_cap, int *free_stack, int *ctl) {
    __shared__ int wsum[32];
    __shared__ int top_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) top_s = 0;
    __syncthreads();
    for (int hi = own_cap - 1; hi >= 0; hi -= 1024) {
        const int g = hi - (int)threadIdx.x;
        const bool fr = (g >= 0) && !S.alive[g];
        const unsigned bal = __ballot_sync(0xffffffffu, fr);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int off = top_s;
        for (int w = 0; w < warp; w++) off += wsum[w];
        if (fr) free_stack[off + __popc(bal & ((1u << lane) - 1u))] = g;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; w++) t += wsum[w];
            top_s += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ctl[0] = top_s;
        ctl[6] = 0;
        ctl[21] = 0;
    }
}

__global__ void count_owned_kernel(DevState S, int own_cap, int *out) {
    int own = 0, ghost = 0;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < S.N; g += gridDim.x * blockDim.x) {
        if (g < own_cap) own += S.alive[g] ? 1 : 0;
        else ghost += (S.pv[g].x == S.pv[g].x) ? 1 : 0;
    }
    own = __reduce_add_sync(0xffffffffu, own);
    ghost = __reduce_add_sync(0xffffffffu, ghost);
    if ((threadIdx.x & 31) == 0) {
        if (own) atomicAdd(out + 0, own);
        if (ghost) atomicAdd(out + 1, ghost);
    }
}

}  // namespace

size_t sbc_domain_record_bytes(void) { return sizeof(DomRecord); }

cudaError_t sbc_domain_scratch_alloc(DomainScratch &D, int own_cap, int max_mig, int max_halo, int n_slots) {
    cudaError_t e;
    if ((e = cudaMalloc((void **)&D.halo_idx, (size_t)n_slots * sizeof(int2))) != cudaSuccess) return e;
    if ((e = cudaMemset(D.halo_idx, 0xff, (size_t)n_slots * sizeof(int2))) != cudaSuccess) return e;
    D.nblk = (own_cap + CLS_PER_BLOCK - 1) / CLS_PER_BLOCK;
    if ((e = cudaMalloc((void **)&D.halo_list, (size_t)2 * (max_mig + max_halo) * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.block_counts, (size_t)D.nblk * 4 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.block_off, (size_t)D.nblk * 4 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.free_stack, (size_t)own_cap * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.ctl, 32 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMemset(D.halo_list, 0, (size_t)2 * (max_mig + max_halo) * sizeof(int))) != cudaSuccess) return e;
    return cudaMemset(D.ctl, 0, 32 * sizeof(int));
}
void sbc_domain_scratch_free(DomainScratch &D) {
    cudaFree(D.block_counts); cudaFree(D.block_off); cudaFree(D.free_stack); cudaFree(D.ctl); cudaFree(D.halo_list); cudaFree(D.halo_idx);
    D = DomainScratch();
}

void sbc_launch_domain_reset(const DevState &S, DomainScratch &D, cudaStream_t st) {
    dom_init_stack_kernel<<<1, 1024, 0, st>>>(S, D.own_cap, D.free_stack, D.ctl);
}

void sbc_launch_domain_pack(const DevParams &P, const DevState &S, DomainScratch &D, void *send_l, void *send_r,
                            cudaStream_t st) {
    const float halo = D.halo;
    dom_count_kernel<<<D.nblk, CLS_THREADS, 0, st>>>(S, D.own_cap, D.x_lo, D.x_hi, P.L, halo, D.block_counts);
    dom_scan_kernel<<<1, 1024, 0, st>>>(D.nblk, D.block_counts, D.block_off, D.ctl, static_cast<DomRecord *>(send_l),
                                        static_cast<DomRecord *>(send_r), D.max_mig, D.max_halo);
    dom_pack_kernel<<<D.nblk, CLS_THREADS, 0, st>>>(S, D.own_cap, D.x_lo, D.x_hi, P.L, halo, D.block_off, D.ctl,
                                                    D.free_stack, static_cast<DomRecord *>(send_l),
                                                    static_cast<DomRecord *>(send_r), D.max_mig, D.max_halo, D.halo_list);
}

void sbc_launch_domain_unpack(const DevState &S, DomainScratch &D, const void *recv_a, const void *recv_b,
                              const void *sent_l, const void *sent_r, int epoch, cudaStream_t st) {
    const int ghost_cap = S.N - D.own_cap;
    const int n = max(ghost_cap, 2 * D.max_mig);
    dom_unpack_kernel<<<(n + 255) / 256, 256, 0, st>>>(S, D.own_cap, static_cast<const DomRecord *>(recv_a),
                                                      static_cast<const DomRecord *>(recv_b),
                                                      static_cast<const DomRecord *>(sent_l),
                                                      static_cast<const DomRecord *>(sent_r), D.max_mig, D.max_halo, D.ctl,
                                                      D.free_stack, D.halo_list, epoch);
    dom_commit_kernel<<<1, 1, 0, st>>>(D.ctl, epoch);
    // the new refresh lists, inverted for the threads that deliver the positions themselves
    cudaMemsetAsync(D.halo_idx, 0xff, (size_t)S.N * sizeof(int2), st);
    const int cap = D.max_halo + D.max_mig;
    dom_halo_idx_kernel<<<(cap + 255) / 256, 256, 0, st>>>(D.ctl, D.halo_list, cap, D.halo_idx);
}

void sbc_launch_domain_refresh_pack(const DevState &S, DomainScratch &D, void *send_l, void *send_r, cudaStream_t st) {
    const int cap = D.max_halo + D.max_mig;
    dom_refresh_pack_kernel<<<(cap + 255) / 256, 256, 0, st>>>(S, D.ctl, D.halo_list, cap, static_cast<DomRecord *>(send_l),
                                                             static_cast<DomRecord *>(send_r));
}
void sbc_launch_domain_refresh_unpack(const DevState &S, DomainScratch &D, const void *recv_a0, const void *recv_a1,
                                      const void *recv_b0, const void *recv_b1, int epoch, cudaStream_t st) {
    const int ghost_cap = S.N - D.own_cap;
    if (ghost_cap <= 0) return;
    dom_refresh_unpack_kernel<<<(ghost_cap + 255) / 256, 256, 0, st>>>(
        S, D.own_cap, D.ctl, static_cast<const DomRecord *>(recv_a0), static_cast<const DomRecord *>(recv_a1),
        static_cast<const DomRecord *>(recv_b0), static_cast<const DomRecord *>(recv_b1), epoch);
}
void sbc_launch_domain_push(DomainScratch &D, const void *send_l, const void *send_r, void *const dst_l[2], void *const dst_r[2],
                            int *const flag_l[2], int *const flag_r[2], int epoch, cudaStream_t st) {
    dom_push_kernel<<<64, 256, 0, st>>>(static_cast<const DomRecord *>(send_l), static_cast<const DomRecord *>(send_r),
                                        static_cast<DomRecord *>(dst_l[0]), static_cast<DomRecord *>(dst_l[1]),
                                        static_cast<DomRecord *>(dst_r[0]), static_cast<DomRecord *>(dst_r[1]), epoch, D.max_mig,
                                        D.max_halo);
    dom_flag_kernel<<<1, 1, 0, st>>>(flag_l[0], flag_l[1], flag_r[0], flag_r[1], epoch);
}
void sbc_launch_domain_refresh_receive(const DevState &S, DomainScratch &D, const uint32_t *step_dev, uint32_t step_off,
                                       const void *const recv_a[2], const void *const recv_b[2], int *const flag_a[2],
                                       int *const flag_b[2], int *const flag_l[2], int *const flag_r[2], cudaStream_t st) {
    const int n_recv = (max(S.N - D.own_cap, 1) + 255) / 256;
    sbc_launch_prio(dom_refresh_receive_kernel, n_recv, 256, 0, st, S, D.own_cap, D.ctl, step_dev, step_off,
                    static_cast<const DomRecord *>(recv_a[0]), static_cast<const DomRecord *>(recv_a[1]),
                    static_cast<const DomRecord *>(recv_b[0]), static_cast<const DomRecord *>(recv_b[1]),
                    (const volatile int *)flag_a[0], (const volatile int *)flag_a[1], (const volatile int *)flag_b[0],
                    (const volatile int *)flag_b[1], D.timeout_ns, (volatile int *)flag_l[0], (volatile int *)flag_l[1],
                    (volatile int *)flag_r[0], (volatile int *)flag_r[1]);
}
// exchange number = (*step_dev + step_off) - 1 + shift, as for the receive: pass the step AFTER the one that delivered
void sbc_launch_domain_publish(DomainScratch &D, const uint32_t *step_dev, uint32_t step_off, int *const flag_l[2],
                               int *const flag_r[2], cudaStream_t st) {
    sbc_launch_prio(dom_publish_kernel, 1, 1, 0, st, (const int *)D.ctl, step_dev, step_off, (volatile int *)flag_l[0],
                    (volatile int *)flag_l[1], (volatile int *)flag_r[0], (volatile int *)flag_r[1]);
}
void sbc_launch_domain_wait(DomainScratch &D, int *const flag_a[2], int *const flag_b[2], int epoch, cudaStream_t st) {
    dom_wait_kernel<<<1, 1, 0, st>>>(flag_a[0], flag_a[1], flag_b[0], flag_b[1], D.ctl, epoch, D.timeout_ns);
}

void sbc_launch_count_owned(const DevState &S, DomainScratch &D, cudaStream_t st) {
    cudaMemsetAsync(D.ctl + 8, 0, 2 * sizeof(int), st);
    count_owned_kernel<<<148, 256, 0, st>>>(S, D.own_cap, D.ctl + 8);
}
