// domain.cu - spatial domain decomposition of ONE large periodic swarm over several handles (one per
// GPU): slabs along x, per-step halo of the agents within att_range of a slab edge, migration of
// agents that crossed an edge (BASELINE config C5). The reference has no
// counterpart (single process); what must be preserved is Step()'s semantics
// (/root/reference/src/swarmdyn.cpp:143-204): interactions see every agent's position at the start of
// the step, every agent is updated once, decisions depend on (seed, GLOBAL agent id, step) only.
//
// Slots of a handle: [0, own_cap) owned agents (alive) or free slots (alive = 0, x = NaN);
// [own_cap, N) ghosts = copies of neighbours' edge agents (alive = 0 but x valid: visible as
// interaction partners j, never rows, never updated). Records travel in fixed-size device buffers
// (header record with the count, then 64-byte records), so no host round trip is needed to size a
// message; the transport itself (NCCL send/recv over NVLink, or a loop-back copy in tests) belongs to
// the host side (sde_burst_coast_b200/domain.py).
//
// Packing order is by slot index (CUB DeviceSelect, stable), arrival order is (left buffer, right
// buffer): results do not depend on thread scheduling.
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "sbc_kernels.cuh"

namespace {

struct __align__(16) DomRecord {  // 64 bytes
    float4 pv;
    float4 hv;
    float2 force;
    uint2 timers;
    uint32_t gid;
    uint32_t pad[3];
};
static_assert(sizeof(DomRecord) == 64, "DomRecord must be 64 bytes");

// which side an agent at x has to go to / be visible from. Slabs are [rank w, (rank+1) w) of a
// periodic box; an owned agent outside its slab belongs to the neighbour it is closer to.
struct HaloPred {
    const float4 *pv;
    const uint8_t *alive;
    float lo, hi;   // select alive owned agents with lo <= x < hi (a band next to one slab edge)
    __device__ bool operator()(int g) const {
        if (!alive[g]) return false;
        const float x = pv[g].x;
        return x >= lo && x < hi;
    }
};
struct LeaverPred {
    const float4 *pv;
    const uint8_t *alive;
    float x_lo, x_hi, L;
    int side;  // 0: leaves through the left edge, 1: through the right edge
    __device__ bool operator()(int g) const {
        if (!alive[g]) return false;
        const float x = pv[g].x;
        if (x >= x_lo && x < x_hi) return false;
        // circular distance to the two edges decides the side
        float dl = x_lo - x; if (dl < 0.f) dl += L;      // how far left of the left edge
        float dr = x - x_hi; if (dr < 0.f) dr += L;      // how far right of the right edge
        return (side == 0) ? (dl <= dr) : (dr < dl);
    }
};
struct FreePred {
    const uint8_t *alive;
    __device__ bool operator()(int g) const { return alive[g] == 0; }
};

__global__ void gather_records_kernel(DevState S, const int *__restrict__ sel, const int *__restrict__ n_sel,
                                      int max_records, int full_state, int free_slots, DomRecord *__restrict__ buf,
                                      int *overflow) {
    const int n = *n_sel;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        DomRecord hdr = {};
        hdr.gid = (uint32_t)min(n, max_records);
        hdr.pad[0] = (uint32_t)n;  // what was wanted; > max_records means overflow
        buf[0] = hdr;
        if (n > max_records) atomicAdd(overflow, 1);
    }
    if (k >= n || k >= max_records) return;
    const int g = sel[k];
    DomRecord r = {};
    r.pv = S.pv[g];
    r.gid = S.gid ? S.gid[g] : (uint32_t)g;
    if (full_state) {
        r.hv = S.hv[g];
        r.force = S.force[g];
        r.timers = S.timers[g];
    }
    buf[1 + k] = r;
    if (free_slots) {  // the agent now lives on the neighbour
        const float qnan = __int_as_float(0x7fc00000);
        S.alive[g] = 0;
        S.pv[g] = make_float4(qnan, qnan, 0.f, 0.f);
    }
}

__global__ void unpack_ghosts_kernel(DevState S, int ghost_first, const DomRecord *__restrict__ a,
                                     const DomRecord *__restrict__ b, int *overflow) {
    const int na = a ? (int)a[0].gid : 0, nb = b ? (int)b[0].gid : 0;
    const int cap = S.N - ghost_first;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0 && na + nb > cap) atomicAdd(overflow, 1);
    if (k >= cap) return;
    const int g = ghost_first + k;
    const float qnan = __int_as_float(0x7fc00000);
    float4 pv = make_float4(qnan, qnan, 0.f, 0.f);
    uint32_t gid = 0xffffffffu;
    if (k < na) { pv = a[1 + k].pv; gid = a[1 + k].gid; }
    else if (k < na + nb) { pv = b[1 + (k - na)].pv; gid = b[1 + (k - na)].gid; }
    S.pv[g] = pv;
    S.alive[g] = 0;
    if (S.gid) S.gid[g] = gid;
}

__global__ void unpack_migrants_kernel(DevState S, const int *__restrict__ free_list, const int *__restrict__ n_free,
                                       const DomRecord *__restrict__ a, const DomRecord *__restrict__ b,
                                       int *overflow) {
    const int na = a ? (int)a[0].gid : 0, nb = b ? (int)b[0].gid : 0;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0 && na + nb > *n_free) atomicAdd(overflow, 1);
    if (k >= na + nb || k >= *n_free) return;
    const DomRecord r = (k < na) ? a[1 + k] : b[1 + (k - na)];
    const int g = free_list[k];
    S.pv[g] = r.pv;
    S.pv_dead[g] = r.pv;
    S.hv[g] = r.hv;
    S.force[g] = r.force;
    S.timers[g] = r.timers;
    S.alive[g] = 1;
    if (S.gid) S.gid[g] = r.gid;
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

template <class Pred>
void select_slots(DomainScratch &D, int n_slots, Pred pred, cudaStream_t st) {
    cub::CountingInputIterator<int> it(0);
    cub::DeviceSelect::If(D.cub_tmp, D.cub_bytes, it, D.sel, D.n_sel, n_slots, pred, st);
}

}  // namespace

size_t sbc_domain_record_bytes(void) { return sizeof(DomRecord); }

cudaError_t sbc_domain_scratch_alloc(DomainScratch &D, int n_slots) {
    cudaError_t e;
    if ((e = cudaMalloc((void **)&D.sel, (size_t)n_slots * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.n_sel, 4 * sizeof(int))) != cudaSuccess) return e;
    cudaMemset(D.n_sel, 0, 4 * sizeof(int));  // [0] selection count, [1..2] owned / ghost counts, [3] overflow events
    cub::CountingInputIterator<int> it(0);
    D.cub_bytes = 0;
    FreePred fp{nullptr};
    cub::DeviceSelect::If(nullptr, D.cub_bytes, it, D.sel, D.n_sel, n_slots, fp);
    D.cub_bytes += 256;
    return cudaMalloc(&D.cub_tmp, D.cub_bytes);
}
void sbc_domain_scratch_free(DomainScratch &D) {
    cudaFree(D.sel); cudaFree(D.n_sel); cudaFree(D.cub_tmp);
    D = DomainScratch();
}

// halo bands: agents within `halo` of the left edge go into buf_left, of the right edge into buf_right
void sbc_launch_pack_halos(const DevParams &P, const DevState &S, DomainScratch &D, void *buf_left,
                           void *buf_right, int max_records, cudaStream_t st) {
    const float halo = P.att * 1.0001f;
    void *bufs[2] = {buf_left, buf_right};
    for (int side = 0; side < 2; side++) {
        if (!bufs[side]) continue;
        HaloPred pr{S.pv, S.alive, side == 0 ? D.x_lo : D.x_hi - halo, side == 0 ? D.x_lo + halo : D.x_hi};
        select_slots(D, D.own_cap, pr, st);
        gather_records_kernel<<<(max_records + 255) / 256, 256, 0, st>>>(S, D.sel, D.n_sel, max_records, 0, 0,
                                                                        static_cast<DomRecord *>(bufs[side]), D.n_sel + 3);
    }
}

void sbc_launch_unpack_ghosts(const DevState &S, DomainScratch &D, const void *a, const void *b, cudaStream_t st) {
    const int cap = S.N - D.own_cap;
    if (cap <= 0) return;
    unpack_ghosts_kernel<<<(cap + 255) / 256, 256, 0, st>>>(S, D.own_cap, static_cast<const DomRecord *>(a),
                                                           static_cast<const DomRecord *>(b), D.n_sel + 3);
}

void sbc_launch_pack_migrants(const DevParams &P, const DevState &S, DomainScratch &D, void *buf_left,
                              void *buf_right, int max_records, cudaStream_t st) {
    void *bufs[2] = {buf_left, buf_right};
    for (int side = 0; side < 2; side++) {
        if (!bufs[side]) continue;
        LeaverPred pr{S.pv, S.alive, D.x_lo, D.x_hi, P.L, side};
        select_slots(D, D.own_cap, pr, st);
        gather_records_kernel<<<(max_records + 255) / 256, 256, 0, st>>>(S, D.sel, D.n_sel, max_records, 1, 1,
                                                                        static_cast<DomRecord *>(bufs[side]), D.n_sel + 3);
    }
}

void sbc_launch_unpack_migrants(const DevState &S, DomainScratch &D, const void *a, const void *b, int max_records,
                                cudaStream_t st) {
    FreePred fp{S.alive};
    select_slots(D, D.own_cap, fp, st);
    unpack_migrants_kernel<<<(2 * max_records + 255) / 256, 256, 0, st>>>(S, D.sel, D.n_sel,
                                                                         static_cast<const DomRecord *>(a),
                                                                         static_cast<const DomRecord *>(b), D.n_sel + 3);
}

void sbc_launch_count_owned(const DevState &S, DomainScratch &D, cudaStream_t st) {
    cudaMemsetAsync(D.n_sel + 1, 0, 2 * sizeof(int), st);
    count_owned_kernel<<<148, 256, 0, st>>>(S, D.own_cap, D.n_sel + 1);
}
