// swarm_block.cu - whole swarms resident on chip: the ensemble path (BASELINE config C2) and every
// single swarm of up to 1024 agents (C1, C3).
//
// The reference's Step() (/root/reference/src/swarmdyn.cpp:143-204) is fused into ONE kernel that
// advances many steps per launch: interaction (InteractionGlobal/IntCalcPrey/SFM_4Zone,
// agents_interact.cpp:144-250) for the rows that start a burst, then ParticleBurstCoast
// (agents_dynamics.cpp:150-273) for every agent. Agent state lives in registers for the whole
// launch (one thread = one agent); global memory is touched at launch entry and exit only.
//
//   swarm_warp_kernel<TPS> : N <= 32. A swarm occupies TPS = 8, 16 or 32 consecutive lanes, 32/TPS
//                            swarms share a warp. Neighbour data moves by warp shuffles, zone
//                            counters by ballot + popc; no shared memory, no block barrier.
//   swarm_block_kernel     : 32 < N <= 1024, one thread block per swarm. Rows that start a burst
//                            are dealt to the warps round-robin; a warp sweeps j over the swarm's
//                            float4 {x,y,vx,vy} tile in shared memory and reduces by shuffles.
// Replicates are independent, so an ensemble is just a larger grid (no communication).
#include "sbc_kernels.cuh"

namespace {

__device__ __forceinline__ void load_agent(const DevState &S, size_t g, AgentState &a) {
    const float4 p = S.pv[g];
    const float4 h = S.hv[g];
    const float2 f = S.force[g];
    const uint2 t = S.timers[g];
    a.x = p.x; a.y = p.y; a.vx = p.z; a.vy = p.w;
    a.phi = h.x; a.vproj = h.y; a.cu = h.z; a.su = h.w;
    a.fx = f.x; a.fy = f.y;
    a.bin_step = t.x; a.stb = t.y;
}
__device__ __forceinline__ void store_agent(const DevState &S, size_t g, const AgentState &a) {
    S.pv[g] = make_float4(a.x, a.y, a.vx, a.vy);
    S.hv[g] = make_float4(a.phi, a.vproj, a.cu, a.su);
    S.force[g] = make_float2(a.fx, a.fy);
    S.timers[g] = make_uint2(a.bin_step, a.stb);
}

__device__ __forceinline__ void draw_decisions(const DevParams &P, const DevState &S, size_t g,
                                               uint32_t rep, uint32_t i, uint32_t step, bool rearm,
                                               double &u_social, float &u_dir, uint32_t &k_next) {
    const uint4 q = sbc_draw(P, rep, i, step, SBC_SLOT_DECISION);
    u_social = S.inj_social ? S.inj_social[g] : (double)q.x * (1.0 / 4294967296.0);
    u_dir = S.inj_dir ? S.inj_dir[g] : (float)(q.y >> 8) * (1.0f / 16777216.0f);
    if (rearm) k_next = S.inj_k ? S.inj_k[g] : sbc_geometric_k(S.geo, P.max_steps, q.z, P.inv_log2q);
}

// ------------------------------------------------------------------------------------------------
// N <= 32 : swarms inside a warp
// ------------------------------------------------------------------------------------------------
constexpr int BC_RUNTIME = -99;   // template value: take the boundary condition from the parameters

template <int TPS, int BCT, bool VOR>
__global__ void __launch_bounds__(128)
swarm_warp_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps) {
    const int BC = (BCT == BC_RUNTIME) ? P.BC : BCT;
    constexpr int SPW = 32 / TPS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int seg = lane / TPS, t = lane % TPS;
    const int rep = warp_global * SPW + seg;
    const int N = S.N;
    const bool valid = (rep < S.R) && (t < N);
    const size_t g = valid ? ((size_t)rep * N + t) : 0;
    const unsigned segmask = (TPS == 32) ? FULL : (((1u << TPS) - 1u) << (seg * TPS));
    AgentState a = {};
    bool alive = false;
    if (valid) {
        load_agent(S, g, a);
        alive = S.alive[g] != 0;
    }
    uint32_t flags_acc = 0;
    unsigned long long n_amb = 0, n_pairs = 0;

    for (uint32_t s = s_first; s != s_first + n_steps; s++) {
        const bool first = alive && (a.bin_step == P.burst_steps);
        RowAcc acc;
        row_acc_clear(acc);
        unsigned rem = __ballot_sync(FULL, first);
        while (rem) {  // warp-uniform: every segment serves its lowest pending row
            const unsigned mine = rem & segmask;
            const bool has = mine != 0;
            const int src = has ? (__ffs(mine) - 1) : lane;
            const float xi = __shfl_sync(FULL, a.x, src), yi = __shfl_sync(FULL, a.y, src);
            int z = 3;
            float ux = 0.f, uy = 0.f;
            if (has && alive && lane != src) {  // pair (row src <- me)
                float dx = a.x - xi, dy = a.y - yi;
                if (BC == 0) {
                    dx = sbc_delta_pbc(xi, a.x, P);
                    dy = sbc_delta_pbc(yi, a.y, P);
                }
                const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                bool amb;
                z = sbc_zone_fast(d2, P, amb);
                if (amb) {
                    z = sbc_zone_exact(xi, yi, a.x, a.y, P);
                    n_amb++;
                }
                const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
                ux = dx * rinv;
                uy = dy * rinv;
                n_pairs++;
            }
            if (VOR) {  // keep only Delaunay neighbours of the row (warp-uniform loop over the tank)
                DelaunayEdge e;
                sbc_delaunay_begin(e, xi, yi, a.x, a.y);
                const unsigned alive_mask = __ballot_sync(FULL, alive);
                const int seg0 = seg * TPS;
                for (int k = 0; k < TPS; k++) {
                    const int lk = seg0 + k;
                    const float xk = __shfl_sync(FULL, a.x, lk), yk = __shfl_sync(FULL, a.y, lk);
                    if (((alive_mask >> lk) & 1u) && lk != src && lk != lane) sbc_delaunay_add(e, xk, yk);
                }
                if (z < 3 && !sbc_delaunay_end(e)) z = 3;
            }
            const int cr = __popc(__ballot_sync(FULL, z == 0) & segmask);
            const int ca = __popc(__ballot_sync(FULL, z == 1) & segmask);
            const int ct = __popc(__ballot_sync(FULL, z == 2) & segmask);
            // the decision reads force_rep when c_rep > 0 and force_alg/force_att otherwise
            // (agents_dynamics.cpp:40-58): reduce only what will be read
            float v0 = (cr > 0) ? ((z == 0) ? -ux : 0.f) : ((z == 1) ? a.vx : 0.f);
            float v1 = (cr > 0) ? ((z == 0) ? -uy : 0.f) : ((z == 1) ? a.vy : 0.f);
            float v2 = (z == 2) ? ux : 0.f;
            float v3 = (z == 2) ? uy : 0.f;
#pragma unroll
            for (int o = TPS / 2; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(FULL, v0, o);
                v1 += __shfl_xor_sync(FULL, v1, o);
                v2 += __shfl_xor_sync(FULL, v2, o);
                v3 += __shfl_xor_sync(FULL, v3, o);
            }
            const bool is_row = has && (lane == src);
            if (is_row) {
                acc.c_rep = cr; acc.c_alg = ca; acc.c_att = ct;
                if (cr > 0) { acc.rep_x = v0; acc.rep_y = v1; }
                else { acc.alg_x = v0; acc.alg_y = v1; }
                acc.att_x = v2; acc.att_y = v3;
            }
            rem &= ~__ballot_sync(FULL, is_row);
        }
        if (alive) {
            double u_social = 0.0;
            float u_dir = 0.f;
            uint32_t k_next = 1;
            const bool rearm = a.stb <= 1;
            if (first || rearm) draw_decisions(P, S, g, rep, t, s, rearm, u_social, u_dir, k_next);
            uint32_t fl = 0;
            sbc_agent_step(a, acc, u_social, u_dir, k_next, P, fl, BC);
            flags_acc |= fl;
        }
    }
    if (valid && alive) {
        store_agent(S, g, a);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
    }
    n_amb = __reduce_add_sync(FULL, (unsigned)n_amb);
    const unsigned np = __reduce_add_sync(FULL, (unsigned)n_pairs);
    if (lane == 0) {
        if (n_amb) atomicAdd(S.stats + 4, n_amb);
        if (np) atomicAdd(S.stats + 3, (unsigned long long)np);
    }
}

// ------------------------------------------------------------------------------------------------
// 32 < N <= 1024 : one block per swarm
// ------------------------------------------------------------------------------------------------
struct RowResult {
    float v0, v1, v2, v3;
    int cr, ca, ct, pad;
};

__global__ void __launch_bounds__(1024)
swarm_block_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    const int N = S.N;
    const int nthreads = blockDim.x;
    const int nwarps = nthreads >> 5;
    float4 *pv_s = reinterpret_cast<float4 *>(smem_raw);                       // [nthreads]
    RowResult *res_s = reinterpret_cast<RowResult *>(pv_s + nthreads);         // [nthreads]
    unsigned *act_s = reinterpret_cast<unsigned *>(res_s + nthreads);          // [32]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int rep = blockIdx.x;
    const bool valid = t < N;
    const size_t g = valid ? ((size_t)rep * N + t) : 0;
    AgentState a = {};
    bool alive = false;
    if (valid) {
        load_agent(S, g, a);
        alive = S.alive[g] != 0;
    }
    uint32_t flags_acc = 0;
    unsigned long long n_amb = 0, n_pairs = 0;
    const float qnan = __int_as_float(0x7fc00000);

    for (uint32_t s = s_first; s != s_first + n_steps; s++) {
        const bool first = alive && (a.bin_step == P.burst_steps);
        RowAcc acc;
        row_acc_clear(acc);
        const int any = __syncthreads_or(first ? 1 : 0);
        if (any) {
            pv_s[t] = alive ? make_float4(a.x, a.y, a.vx, a.vy) : make_float4(qnan, qnan, 0.f, 0.f);
            const unsigned m = __ballot_sync(FULL, first);
            if (lane == 0) act_s[warp] = m;
            __syncthreads();
            int k = 0;
            for (int ww = 0; ww < nwarps; ww++) {
                unsigned m2 = act_s[ww];
                while (m2) {
                    const int bit = __ffs(m2) - 1;
                    m2 &= m2 - 1;
                    if ((k++ % nwarps) != warp) continue;
                    const int row = ww * 32 + bit;
                    const float4 pi = pv_s[row];
                    int cr = 0, ca = 0, ct = 0;
                    float rx = 0.f, ry = 0.f, gx = 0.f, gy = 0.f, tx = 0.f, ty = 0.f;
                    for (int j = lane; j < N; j += 32) {
                        const float4 pj = pv_s[j];
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
                        if (j == row) z = 3;
                        if (P.voronoi && z < 3) {  // Delaunay edge test against every other agent of the swarm
                            DelaunayEdge e;
                            sbc_delaunay_begin(e, pi.x, pi.y, pj.x, pj.y);
                            for (int k = 0; k < N; k++) {
                                const float4 pk = pv_s[k];
                                if (k != row && k != j && pk.x == pk.x) sbc_delaunay_add(e, pk.x, pk.y);
                            }
                            if (!sbc_delaunay_end(e)) z = 3;
                        }
                        const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
                        if (z == 0) { rx = __fmaf_rn(-dx, rinv, rx); ry = __fmaf_rn(-dy, rinv, ry); cr++; }
                        else if (z == 1) { gx += pj.z; gy += pj.w; ca++; }
                        else if (z == 2) { tx = __fmaf_rn(dx, rinv, tx); ty = __fmaf_rn(dy, rinv, ty); ct++; }
                    }
                    n_pairs += (lane == 0) ? (unsigned long long)(N - 1) : 0ull;
                    cr = __reduce_add_sync(FULL, cr);
                    ca = __reduce_add_sync(FULL, ca);
                    ct = __reduce_add_sync(FULL, ct);
                    float v0 = (cr > 0) ? rx : gx, v1 = (cr > 0) ? ry : gy, v2 = tx, v3 = ty;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        v0 += __shfl_xor_sync(FULL, v0, o);
                        v1 += __shfl_xor_sync(FULL, v1, o);
                        v2 += __shfl_xor_sync(FULL, v2, o);
                        v3 += __shfl_xor_sync(FULL, v3, o);
                    }
                    if (lane == 0) {
                        RowResult r;
                        r.v0 = v0; r.v1 = v1; r.v2 = v2; r.v3 = v3;
                        r.cr = cr; r.ca = ca; r.ct = ct; r.pad = 0;
                        res_s[row] = r;
                    }
                }
            }
            __syncthreads();
            if (first) {
                const RowResult r = res_s[t];
                acc.c_rep = r.cr; acc.c_alg = r.ca; acc.c_att = r.ct;
                if (r.cr > 0) { acc.rep_x = r.v0; acc.rep_y = r.v1; }
                else { acc.alg_x = r.v0; acc.alg_y = r.v1; }
                acc.att_x = r.v2; acc.att_y = r.v3;
            }
        }
        if (alive) {
            double u_social = 0.0;
            float u_dir = 0.f;
            uint32_t k_next = 1;
            const bool rearm = a.stb <= 1;
            if (first || rearm) draw_decisions(P, S, g, rep, t, s, rearm, u_social, u_dir, k_next);
            uint32_t fl = 0;
            sbc_agent_step(a, acc, u_social, u_dir, k_next, P, fl, P.BC);
            flags_acc |= fl;
        }
    }
    if (valid && alive) {
        store_agent(S, g, a);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
    }
    if (n_amb) atomicAdd(S.stats + 4, n_amb);
    if (n_pairs) atomicAdd(S.stats + 3, n_pairs);
}

}  // namespace

bool sbc_launch_swarm_block(const DevParams &P, const DevState &S, long long s_first,
                            long long n_steps, cudaStream_t st) {
    const int N = S.N;
    if (N > 1024 || N < 1) return false;
    const uint32_t s0 = (uint32_t)s_first, ns = (uint32_t)n_steps;
    if (N <= 32) {
        const int tps = (N <= 8) ? 8 : (N <= 16) ? 16 : 32;
        const int spw = 32 / tps;
        const long long warps = ((long long)S.R + spw - 1) / spw;
        const int wpb = 4;
        const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
#define SBC_WARP_LAUNCH(T, B)                                                                   \
    do {                                                                                        \
        if (P.voronoi) swarm_warp_kernel<T, B, true><<<blocks, wpb * 32, 0, st>>>(P, S, s0, ns); \
        else swarm_warp_kernel<T, B, false><<<blocks, wpb * 32, 0, st>>>(P, S, s0, ns);         \
    } while (0)
#define SBC_WARP_BC(T)                                                  \
    do {                                                                \
        if (P.BC == 6) SBC_WARP_LAUNCH(T, 6);                           \
        else if (P.BC == 0) SBC_WARP_LAUNCH(T, 0);                      \
        else SBC_WARP_LAUNCH(T, BC_RUNTIME);                            \
    } while (0)
        if (tps == 8) SBC_WARP_BC(8);
        else if (tps == 16) SBC_WARP_BC(16);
        else SBC_WARP_BC(32);
#undef SBC_WARP_BC
#undef SBC_WARP_LAUNCH
        return true;
    }
    const int nthreads = ((N + 31) / 32) * 32;
    const size_t smem = (size_t)nthreads * (sizeof(float4) + sizeof(RowResult)) + 32 * sizeof(unsigned);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(swarm_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    swarm_block_kernel<<<S.R, nthreads, smem, st>>>(P, S, s0, ns);
    return true;
}
