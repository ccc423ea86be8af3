// sbc_device.cuh - device-side building blocks shared by every kernel of libsbc.so (sm_100a).
//
// FP32 structure-of-arrays state, FP64 only in the rare exact re-checks that make integer
// results (zone ids, neighbour counts, burst timers) bit-identical to the reference.
// Reference lines are cited as file:line into /root/reference/src.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SBC_PI_F 3.14159265358979323846f
#define SBC_TWO_PI_F 6.28318530717958647692f
#define SBC_PI_D 3.14159265358979323846

// ---------------------------------------------------------------------------------------------
// parameters as the kernels see them
// ---------------------------------------------------------------------------------------------
struct DevParams {
    // geometry / dynamics (FP32 copies of sbc_params)
    float L, invL, L_lo;   // L_lo = sizeL - (double)(float)sizeL
    float dt, beta, alphaTurn;
    float soc, env, inv_soc, inv_env;
    float inv_log2q;               // 1 / log2(1 - burst_rate dt): continuous inverse CDF of the inter-burst draw
    float rep, alg, att;           // ranges
    float rep2, alg2, att2;        // squared thresholds for the fast compare
    float band_lo[3], band_hi[3];  // d2 inside [lo,hi] of any threshold => FP64 re-check
    float far2;                    // dense pass: d2 above this is certainly outside att_range
    float compact_lim;             // dense pass, periodic: max |block-relative coordinate| of a row
    float flee_range, inv_flee_range;
    float pred_speed, kill_range, kill_rate, noisep;
    float wall_margin_r;           // sizeL - 2 (agents_dynamics.cpp:111)
    float t_burst;                 // dt * burst_steps
    // exact copies for FP64 re-checks
    double dL, d_rep, d_alg, d_att, d_prob_social, d_flee_range, d_kill_range;
    // integers
    uint32_t burst_steps, max_steps;
    uint32_t tm_bits, tm_mask, tm_stb_max;   // packed burst timers: bin_step in the low tm_bits, steps_till_burst above
    int32_t BC, N_confu, pred_kill, pred_move;
    int32_t has_alg_zone;          // alg_range > rep_range
    int32_t voronoi;               // neighbour candidates = Delaunay edges (InteractionVoronoiF2F) instead of all agents
    // Philox key / ids
    uint32_t seed_lo, seed_hi;
    uint32_t replicate_offset;
};

// per-agent persistent state (registers in the block kernel, SoA arrays in global memory)
struct AgentState {
    float x, y, vx, vy;
    float phi, vproj;
    float cu, su;   // cached (cos, sin)(phi), carried in the state so that launches can be split anywhere
    float fx, fy;
    uint32_t bin_step, stb;
};

// burst timers of one agent in one word (agents.h:50-51: bin_step <= burst_steps, steps_till_burst <= max_steps + 1;
// the split is chosen at sbc_create so that both fit)
__device__ __forceinline__ uint32_t sbc_tm_pack(uint32_t bin_step, uint32_t stb, const DevParams &P) {
    return bin_step | (min(stb, P.tm_stb_max) << P.tm_bits);
}
__device__ __forceinline__ uint32_t sbc_tm_bin(uint32_t t, const DevParams &P) { return t & P.tm_mask; }
__device__ __forceinline__ uint32_t sbc_tm_stb(uint32_t t, const DevParams &P) { return t >> P.tm_bits; }

// accumulators of one row (agents_interact.cpp:211-225, :286-290)
struct RowAcc {
    float rep_x, rep_y, alg_x, alg_y, att_x, att_y, flee_x, flee_y;
    int c_rep, c_alg, c_att, c_flee;
};
__device__ __forceinline__ void row_acc_clear(RowAcc &a) {
    a.rep_x = a.rep_y = a.alg_x = a.alg_y = a.att_x = a.att_y = a.flee_x = a.flee_y = 0.f;
    a.c_rep = a.c_alg = a.c_att = a.c_flee = 0;
}

// branch flags, same bit meaning as orc_step() in oracle/sbc_oracle.h
enum {
    SBC_FL_BURST_START = 1, SBC_FL_SOCIAL = 2, SBC_FL_WALL = 4, SBC_FL_VPROJ_HACK = 8,
    SBC_FL_OVERSHOOT = 16, SBC_FL_BOUNDARY = 32, SBC_FL_FLEE = 64, SBC_FL_REARM = 128
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter (agent, step, slot, replicate), key (seed_lo, seed_hi)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint4 sbc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                            uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

#define SBC_SLOT_DECISION 0u
#define SBC_SLOT_DETECT0 1u
#define SBC_PREDATOR_AGENT_ID 0xFFFFFFFFu

__device__ __forceinline__ uint4 sbc_draw(const DevParams &P, uint32_t replicate, uint32_t agent,
                                          uint32_t step, uint32_t slot) {
    return sbc_philox4x32_10(agent, step, slot, replicate + P.replicate_offset, P.seed_lo, P.seed_hi);
}

// smallest k >= 1 with r < thr[k-1]; none -> max_steps + 1 (inverse CDF of the rejection loop
// agents_dynamics.cpp:250-270, integer-exact against the oracle's table). The continuous inverse
// CDF gives the position in the table to within a step or two; the table itself decides.
__device__ __forceinline__ uint32_t sbc_geometric_k(const uint32_t *__restrict__ thr, uint32_t n,
                                                    uint32_t r, float inv_log2q) {
    const float one_minus_u = fmaxf((float)(~r) * (1.0f / 4294967296.0f), 1e-12f);
    const float kf = __log2f(one_minus_u) * inv_log2q;
    uint32_t lo = (uint32_t)fminf(fmaxf(kf, 0.f), (float)n);   // ~ #{j : thr[j] <= r}
    while (lo < n && __ldg(thr + lo) <= r) lo++;
    while (lo > 0 && __ldg(thr + lo - 1) > r) lo--;
    return lo + 1;
}

__device__ __forceinline__ float sbc_rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
// Minimum-image displacement x_j - x_i for positions kept in [0, L) by Boundary case 0, so one
// fold is enough. Agrees with fmod(r + sgn(r) L/2, L) - sgn(r) L/2 (mathtools.h:64-71) away from
// the |r| = L/2 tie, which lies outside att_range (asserted at sbc_create).
//   rough  : r - L rint(r/L); error up to ulp(L) because x_j - x_i itself rounds when the pair
//            straddles the box edge. Good enough to decide "far".
//   precise: the coordinate on the far side of the edge is shifted by L first - exact by
//            Sterbenz since it lies in [L/2, L) - so the final subtraction is of two nearby
//            numbers and the result is correct to ulp(|dx|), like an unwrapped pair. The low
//            word of a box length that is not FP32-representable is applied last.
__device__ __forceinline__ float sbc_fold_count(float r, const DevParams &P) {
    // round-to-nearest of r/L through the 1.5*2^23 magic constant: FFMA + FADD, no FRND
    return __fadd_rn(__fmaf_rn(r, P.invL, 12582912.f), -12582912.f);
}
__device__ __forceinline__ float sbc_delta_pbc_rough(float xi, float xj, const DevParams &P) {
    const float r = xj - xi;
    return __fmaf_rn(-P.L, sbc_fold_count(r, P), r);
}
__device__ __forceinline__ float sbc_delta_pbc(float xi, float xj, const DevParams &P) {
    const float n = sbc_fold_count(xj - xi, P);            // -1, 0 or +1
    const float a = __fmaf_rn(-fmaxf(n, 0.f), P.L, xj);   // n = +1: x_j - L
    const float b = __fmaf_rn(fminf(n, 0.f), P.L, xi);    // n = -1: x_i - L  (so a - b = r + L)
    return __fmaf_rn(-n, P.L_lo, a - b);
}

// FP64 minimum image exactly as mathtools.h:64-71 computes it: fmod(r + sgn(r) L/2, L) - sgn(r) L/2.
// For |t| < 2 L the fmod is one exact subtraction (Sterbenz), so the slow library path is only kept
// for arguments the simulation never produces.
__device__ __forceinline__ double sbc_fmod_near(double t, double L) {
    const double at = fabs(t);
    if (at < L) return t;
    if (at < 2.0 * L) return (t > 0.0) ? __dsub_rn(t, L) : __dadd_rn(t, L);
    return fmod(t, L);
}
__device__ __forceinline__ double sbc_min_image_dd(double r, double L) {
    const double s = (double)((0.0 < r) - (r < 0.0));
    const double h = __ddiv_rn(__dmul_rn(s, L), 2.0);
    return __dsub_rn(sbc_fmod_near(__dadd_rn(r, h), L), h);
}

// FP64 replay of CalcDistVec + vec_length + SFM_4Zone in the reference's operation order
// (mathtools.h:54-83, social_forces.cpp:33-47): returns 0,1,2 or 3 (= none).
static __device__ __noinline__ int sbc_zone_exact(float xi, float yi, float xj, float yj,
                                           const DevParams &P) {
    double r[2] = {__dsub_rn((double)xj, (double)xi), __dsub_rn((double)yj, (double)yi)};
    if (P.BC == 0) {
        // (rare path, kept in this exact form: the register allocation of the ensemble kernel around the
        // call changes with the body of this function - 6 % of bench.py's headline figure)
#pragma unroll
        for (int d = 0; d < 2; d++) {
            const double s = (double)((0.0 < r[d]) - (r[d] < 0.0));
            const double h = __ddiv_rn(__dmul_rn(s, P.dL), 2.0);
            const double t = fmod(__dadd_rn(r[d], h), P.dL);
            r[d] = __dsub_rn(t, h);
        }
    }
    const double len = __dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1]));
    const double d = __dsqrt_rn(len);
    if (d <= P.d_rep) return 0;
    if (d > P.d_rep && d <= P.d_alg) return 1;
    if (d > P.d_alg && d <= P.d_att) return 2;
    return 3;
}

// Sum the zone accumulators of one row over a thread block in a fixed order (lanes by xor tree, then
// warps in index order): the result does not depend on scheduling. `sh` holds 9 words per warp.
// Returns the totals in thread 0 (other threads get partial data).
template <int NWARPS>
__device__ __forceinline__ void sbc_block_reduce_row(RowAcc &A, float (*shf)[6], int (*shi)[3]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        A.rep_x += __shfl_xor_sync(0xffffffffu, A.rep_x, o);
        A.rep_y += __shfl_xor_sync(0xffffffffu, A.rep_y, o);
        A.alg_x += __shfl_xor_sync(0xffffffffu, A.alg_x, o);
        A.alg_y += __shfl_xor_sync(0xffffffffu, A.alg_y, o);
        A.att_x += __shfl_xor_sync(0xffffffffu, A.att_x, o);
        A.att_y += __shfl_xor_sync(0xffffffffu, A.att_y, o);
    }
    A.c_rep = __reduce_add_sync(0xffffffffu, A.c_rep);
    A.c_alg = __reduce_add_sync(0xffffffffu, A.c_alg);
    A.c_att = __reduce_add_sync(0xffffffffu, A.c_att);
    __syncthreads();  // previous row's readers are done with the buffers
    if (lane == 0) {
        shf[warp][0] = A.rep_x; shf[warp][1] = A.rep_y; shf[warp][2] = A.alg_x;
        shf[warp][3] = A.alg_y; shf[warp][4] = A.att_x; shf[warp][5] = A.att_y;
        shi[warp][0] = A.c_rep; shi[warp][1] = A.c_alg; shi[warp][2] = A.c_att;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NWARPS; w++) {
            A.rep_x += shf[w][0]; A.rep_y += shf[w][1]; A.alg_x += shf[w][2];
            A.alg_y += shf[w][3]; A.att_x += shf[w][4]; A.att_y += shf[w][5];
            A.c_rep += shi[w][0]; A.c_alg += shi[w][1]; A.c_att += shi[w][2];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Voronoi neighbour mode (InteractionVoronoiF2F, agents_interact.cpp:26-77): j is a candidate of row i
// iff (i, j) is an edge of the Delaunay triangulation of the alive agents. Edge test without building
// a triangulation: a circle through p_i and p_j with centre m + t n on their perpendicular bisector
// is empty iff t lies in an interval that every other agent k narrows from one side; the edge exists
// iff the interval stays non-empty. Evaluated in FP64 with explicitly rounded operations in the
// order of the CPU oracles, so the decisions are identical to theirs (min / max over k do not
// depend on the order in which the k arrive).
// ---------------------------------------------------------------------------------------------
struct DelaunayEdge {
    double mx, my, nx, ny, q4, lo, hi;
    bool ok;
};
// Periodic boxes (InteractionVoronoiF2FP with GetCopies4PeriodicBC, agents_operation.cpp:240-293): the reference means
// to triangulate the agents together with shifted copies that give every agent its surroundings up to L/2 in each
// direction (its quadrant flags are never reset - undefined behaviour). The well-defined equivalent used here:
// the test for row i sees every other point at its image NEAREST to row i, which is the periodic Delaunay
// triangulation as long as the empty circles are smaller than L/2. Predators are vertices of the reference's
// triangulation in the predator phase (agents_interact.cpp:103-110): they are added as points that can only remove edges.
__device__ __forceinline__ double sbc_vor_pos(float ref, float v, const DevParams &P) {
    if (P.BC != 0) return (double)v;
    return __dadd_rn((double)ref, sbc_min_image_dd(__dsub_rn((double)v, (double)ref), P.dL));
}
__device__ __forceinline__ void sbc_delaunay_begin(DelaunayEdge &e, double xi, double yi, double xj, double yj) {
    const double ax = xi, ay = yi, bx = xj, by = yj;
    const double dx = __dsub_rn(bx, ax), dy = __dsub_rn(by, ay);
    const double len2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    e.mx = __dmul_rn(0.5, __dadd_rn(ax, bx));
    e.my = __dmul_rn(0.5, __dadd_rn(ay, by));
    e.nx = -dy;
    e.ny = dx;
    e.q4 = __dmul_rn(0.25, len2);
    e.lo = -INFINITY;
    e.hi = INFINITY;
    e.ok = (len2 != 0.0);
}
__device__ __forceinline__ void sbc_delaunay_add(DelaunayEdge &e, double xk, double yk) {
    const double qx = __dsub_rn(xk, e.mx), qy = __dsub_rn(yk, e.my);
    const double c = __dsub_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), e.q4);
    const double s = __dadd_rn(__dmul_rn(qx, e.nx), __dmul_rn(qy, e.ny));
    if (s > 0.0) {
        e.hi = fmin(e.hi, __ddiv_rn(c, __dmul_rn(2.0, s)));
    } else if (s < 0.0) {
        e.lo = fmax(e.lo, __ddiv_rn(c, __dmul_rn(2.0, s)));
    } else if (c < 0.0) {
        e.ok = false;
    }
}
__device__ __forceinline__ bool sbc_delaunay_end(const DelaunayEdge &e) { return e.ok && !(e.lo > e.hi); }

// fast zone from the FP32 squared distance; `amb` is set when d2 falls inside the rounding band
// of any threshold (then the caller must ask sbc_zone_exact)
__device__ __forceinline__ int sbc_zone_fast(float d2, const DevParams &P, bool &amb) {
    amb = (d2 >= P.band_lo[0] && d2 <= P.band_hi[0]) || (d2 >= P.band_lo[1] && d2 <= P.band_hi[1]) ||
          (d2 >= P.band_lo[2] && d2 <= P.band_hi[2]);
    return (d2 <= P.rep2) ? 0 : (d2 <= P.alg2) ? 1 : (d2 <= P.att2) ? 2 : 3;
}

// prey <- predator detection, agents_interact.cpp:252-292; u_detect from the agent's own draw
__device__ __forceinline__ void sbc_detect_accumulate(RowAcc &acc, float xi, float yi, float px,
                                                      float py, double u_detect,
                                                      const DevParams &P) {
    float dx = px - xi, dy = py - yi;
    if (P.BC == 0) {
        dx = sbc_delta_pbc(xi, px, P);
        dy = sbc_delta_pbc(yi, py, P);
    }
    const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
    const float d = sqrtf(d2);
    const float fs = 2.f / (1.f + d * P.inv_flee_range);
    bool hit = (float)u_detect < fs;
    // the decision `u < 2/(1+d/flee_range)` is re-taken in FP64 when FP32 cannot separate the two
    if (fabsf((float)u_detect - fs) <= 4e-6f * fmaxf(fs, 1.f)) {
        double r[2] = {__dsub_rn((double)px, (double)xi), __dsub_rn((double)py, (double)yi)};
        if (P.BC == 0) {
            r[0] = sbc_min_image_dd(r[0], P.dL);
            r[1] = sbc_min_image_dd(r[1], P.dL);
        }
        const double dd = __dsqrt_rn(__dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1])));
        const double fsd = __ddiv_rn(2.0, __dadd_rn(1.0, __ddiv_rn(dd, P.d_flee_range)));
        hit = u_detect < fsd;
    }
    if (hit) {
        const float rinv = (d2 > 0.f) ? rsqrtf(d2) : 0.f;
        acc.c_flee++;
        acc.flee_x -= dx * rinv;
        acc.flee_y -= dy * rinv;
    }
}

// ---------------------------------------------------------------------------------------------
// PredKill / FishNetKill predicates (agents_dynamics.cpp:780-915), shared by every kernel that kills.
// The comparisons run in FP64 in the reference's operation order, but only for agents an FP32 estimate
// cannot rule out: the estimate's error (a few ulp of L) is far below the margin used here.
// ---------------------------------------------------------------------------------------------
// distance prey - predator in FP64, the reference's operation order
__device__ __forceinline__ double sbc_kill_distance(float ax, float ay, float px, float py, const DevParams &P) {
    double r[2] = {__dsub_rn((double)ax, (double)px), __dsub_rn((double)ay, (double)py)};
    if (P.BC == 0) {
        r[0] = sbc_min_image_dd(r[0], P.dL);
        r[1] = sbc_min_image_dd(r[1], P.dL);
    }
    return __dsqrt_rn(__dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1])));
}
// where a prey stands relative to the predator: sensed (d <= 4 kill_range, :861-870), candidate
// (d < kill_range, :872) and within (d <= kill_range, the radial kill :845-850). Decided on the FP32
// squared distance; only a distance within a few ulp of a threshold is re-decided in FP64.
struct KillClass {
    bool sensed, cand, within;
};
__device__ __forceinline__ KillClass sbc_kill_classify(float ax, float ay, float px, float py, const DevParams &P) {
    float dx = ax - px, dy = ay - py;
    if (P.BC == 0) {
        dx = sbc_delta_pbc(px, ax, P);
        dy = sbc_delta_pbc(py, ay, P);
    }
    const float d2 = __fmaf_rn(dy, dy, dx * dx);
    const float kr2 = P.kill_range * P.kill_range, rs2 = 16.f * kr2;
    KillClass k;
    k.sensed = d2 <= rs2;
    k.cand = d2 < kr2;
    k.within = d2 <= kr2;
    if (fabsf(d2 - kr2) <= 8e-6f * kr2 || fabsf(d2 - rs2) <= 8e-6f * rs2) {
        const double d = sbc_kill_distance(ax, ay, px, py, P);
        k.sensed = d <= 4 * P.d_kill_range;
        k.cand = k.sensed && d < P.d_kill_range;
        k.within = d <= P.d_kill_range;
    }
    return k;
}
// kill probability of a candidate (d < kill_range, sensed), :879-905. eff_table[n] is the confusion-
// modulated rate kill_rate / 2 (tanh(-(n - N_confu)) + 1) for n sensed prey, tabulated by the host in FP64.
__device__ __forceinline__ double sbc_kill_probability(float ax, float ay, float px, float py, int n_sensed,
                                                       double total, const double *eff_table, const DevParams &P) {
    if (P.pred_kill == 2) return (double)P.kill_rate * (double)P.dt;
    const double eff = eff_table[n_sensed];
    if (P.pred_kill == 3) return eff * (double)P.dt;
    if (P.pred_kill == 4) {
        const double d = sbc_kill_distance(ax, ay, px, py, P);
        return eff * (double)P.dt * (((P.d_kill_range - d) / P.d_kill_range) / total);
    }
    return 0.0;
}
// selection weight of a candidate for kill rule 4, :872-877
__device__ __forceinline__ double sbc_kill_weight(float ax, float ay, float px, float py, const DevParams &P) {
    return (P.d_kill_range - sbc_kill_distance(ax, ay, px, py, P)) / P.d_kill_range;
}
// FishNetKill :780-820 for one agent: the net's first two nodes p0, p1, heading phi0 of node 0
__device__ __forceinline__ bool sbc_net_kills(float ax, float ay, float4 p0, float4 p1, float phi0, int NP,
                                              const DevParams &P) {
    float s0, c0;
    sincosf(phi0, &s0, &c0);
    {   // FP32 estimate of the distance in front of the net
        float rx = ax - p0.x, ry = ay - p0.y;
        if (P.BC == 0) {
            rx = sbc_delta_pbc(p0.x, ax, P);
            ry = sbc_delta_pbc(p0.y, ay, P);
        }
        const float front = __fmaf_rn(c0, rx, s0 * ry);
        const float slack = __fmaf_rn(0.01f, P.kill_range, 1e-5f * P.L);
        if (front < -slack || front > P.kill_range + slack) return false;
    }
    double vn[2] = {__dsub_rn((double)p1.x, (double)p0.x), __dsub_rn((double)p1.y, (double)p0.y)};
    if (P.BC == 0) { vn[0] = sbc_min_image_dd(vn[0], P.dL); vn[1] = sbc_min_image_dd(vn[1], P.dL); }
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(vn[0], vn[0]), __dmul_rn(vn[1], vn[1])));
    const double net_len = (double)(NP - 1) * dist;
    double r[2] = {__dsub_rn((double)ax, (double)p0.x), __dsub_rn((double)ay, (double)p0.y)};
    if (P.BC == 0) { r[0] = sbc_min_image_dd(r[0], P.dL); r[1] = sbc_min_image_dd(r[1], P.dL); }
    const double front = __dadd_rn(__dmul_rn((double)c0, r[0]), __dmul_rn((double)s0, r[1]));
    const double side = __dadd_rn(__dmul_rn(vn[0], r[0]), __dmul_rn(vn[1], r[1]));
    return front > 0 && front < P.d_kill_range && side > 0 && side <= net_len;
}

// ---------------------------------------------------------------------------------------------
// wall look-ahead (agents_dynamics.cpp:627-772) in closed form.
//
// With constant force F during tb = min(t_burst, t_tnb) and coasting for t_tnb - tb afterwards,
// predictXatNextBurst is LINEAR in v and F:  x_fut = x + A v + B F, with
//   E1 = exp(-beta tb), E2 = exp(-beta (t_tnb - tb)),
//   A = (1 - E1)/beta + E1 (1 - E2)/beta,
//   B = tb/beta - (1 - E1)/beta^2 + (1 - E1)(1 - E2)/beta^2
// (expand predictPos(v,F,tb) + predictPos(predictV(v,F,tb),0,t_tnb-tb)). A and B do not depend on
// the force direction, so the direction search of forceChange2NotCollide only re-evaluates
// |P + R (cos a, sin a)| with P = x + A v and R = B |F| - no exponentials inside the search.
// ---------------------------------------------------------------------------------------------
struct WallCoef { float A, B; };
__device__ __forceinline__ WallCoef sbc_wall_coef(uint32_t stb, const DevParams &P) {
    const float t_tnb = P.dt * (float)stb;
    const float tb = fminf(P.t_burst, t_tnb);
    const float E1 = expf(-P.beta * tb), E2 = expf(-P.beta * (t_tnb - tb));
    const float ib = 1.f / P.beta;
    WallCoef c;
    c.A = (1.f - E1) * ib + E1 * (1.f - E2) * ib;
    c.B = tb * ib - (1.f - E1) * ib * ib + (1.f - E1) * (1.f - E2) * ib * ib;
    return c;
}

// forceChange2NotCollide, agents_dynamics.cpp:699-756 (same search schedule, cheap predictor)
__device__ __forceinline__ float sbc_force_change(float px, float py, float R, float fang,
                                                  float dphi, const DevParams &P) {
    float inside_change = SBC_PI_F;
    bool outside = true, found = false;
    float change = dphi;
    const float L2 = P.L * P.L;
    // The reference loop ends at the first inside probe once |dphi| <= pi/64. In FP32 the inside test
    // is not monotone at ulp scale, so a run of outside probes could shrink dphi below ulp(change)
    // and never end: the iteration count is bounded (4 coarse probes + 24 halvings reach 1e-8 rad)
    // and the last inside angle is returned.
    for (int it = 0; it < 40 && fabsf(change) < SBC_PI_F && outside; it++) {
        float s, c;
        sincosf(fang + change, &s, &c);
        const float qx = px + R * c, qy = py + R * s;
        if (qx * qx + qy * qy < L2) {
            inside_change = change;
            found = true;
            outside = false;
        }
        if (found) dphi *= 0.5f;
        if (outside) {
            change += dphi;
        } else if (fabsf(dphi) > SBC_PI_F / 64.f) {
            change -= dphi;
            outside = true;
        }
    }
    return inside_change;
}

// ---------------------------------------------------------------------------------------------
// burst decision, agents_dynamics.cpp:23-121 (draw_social_or_environmental_force)
// returns the force (magnitude force_mag) and updates flags
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sbc_decide(const AgentState &a, const RowAcc &acc, double u_social,
                                           float u_dir, const DevParams &P, float &fx, float &fy,
                                           float &force_mag, uint32_t &fl, const int BC) {
    float gx = 0.f, gy = 0.f;
    force_mag = P.soc;
    if (u_social <= P.d_prob_social) {  // :35, taken in FP64: u = r * 2^-32 is exact
        fl |= SBC_FL_SOCIAL;
        if (acc.c_rep > 0) {  // :40
            gx = acc.rep_x;
            gy = acc.rep_y;
        } else {
            if (acc.alg_x != 0.f || acc.alg_y != 0.f) {  // :47-52 set_mag(force_alg, c_alg)
                const float s = (float)acc.c_alg * rsqrtf(acc.alg_x * acc.alg_x + acc.alg_y * acc.alg_y);
                gx += acc.alg_x * s;
                gy += acc.alg_y * s;
            }
            if (acc.att_x != 0.f || acc.att_y != 0.f) {  // :53-58
                const float s = (float)acc.c_att * rsqrtf(acc.att_x * acc.att_x + acc.att_y * acc.att_y);
                gx += acc.att_x * s;
                gy += acc.att_y * s;
            }
        }
        if (gx != 0.f || gy != 0.f) {  // :62-65
            const float s = rsqrtf(gx * gx + gy * gy);
            gx *= s;
            gy *= s;
        } else {  // :67-71
            sincosf(a.phi, &gy, &gx);
        }
    } else {
        if (acc.c_flee > 0) {  // :76-80 (a zero flee sum gives NaN in the reference as well)
            fl |= SBC_FL_FLEE;
            const float s = rsqrtf(acc.flee_x * acc.flee_x + acc.flee_y * acc.flee_y);
            gx = acc.flee_x * s;
            gy = acc.flee_y * s;
        } else {  // :81-89
            force_mag = P.env;
            sincosf(SBC_TWO_PI_F * u_dir, &gy, &gx);
        }
    }
    if (BC >= 5) {  // :93-117
        const WallCoef w = sbc_wall_coef(a.stb, P);
        const float px = a.x + w.A * a.vx, py = a.y + w.A * a.vy;
        const float R = w.B * force_mag;
        const float qx = px + R * gx, qy = py + R * gy;
        if (qx * qx + qy * qy > P.wall_margin_r * P.wall_margin_r) {
            fl |= SBC_FL_WALL;
            const float fang = atan2f(gy, gx);
            const float ccw = sbc_force_change(px, py, R, fang, SBC_PI_F / 4.f, P);
            const float cw = sbc_force_change(px, py, R, fang, -SBC_PI_F / 4.f, P);
            float ang = ccw;
            if (fabsf(cw) < fabsf(ang)) ang = cw;
            sincosf(ang + fang, &gy, &gx);
        }
    }
    fx = force_mag * gx;
    fy = force_mag * gy;
}

// ---------------------------------------------------------------------------------------------
// Boundary<agent>, agents_dynamics.cpp:276-471. Returns true when the state was changed.
// BC 0 is applied as a conditional shift (== fmod(x + L, L) for x in [-L, 2L), without the
// rounding that x + L would cost in FP32).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sbc_boundary(float &x, float &y, float &vx, float &vy, float &phi,
                                             const DevParams &P, const int BC) {
    const float L = P.L;
    switch (BC) {
    case 0: {
        if (x < 0.f) x += L; else if (x >= L) x -= L;
        if (y < 0.f) y += L; else if (y >= L) y -= L;
        return false;
    }
    case 5:
    case 6: {
        const float r2 = x * x + y * y;
        if (r2 > L * L) {
            const float r = sqrtf(r2);
            const float nx = x / r, ny = y / r;
            const float diff = r - L;
            x -= 2.f * diff * nx;
            y -= 2.f * diff * ny;
            const float vn = nx * vx + ny * vy;
            const float g = (BC == 5) ? 2.f * vn : vn;
            vx -= g * nx;
            vy -= g * ny;
            phi = atan2f(vy, vx);  // u = v/|v|, phi = atan2(u) (:447-449, :467-469)
            return true;
        }
        return false;
    }
    case 1: {  // inelastic box :300-356
        const float dphi = 0.1f;
        bool hit = false;
        float t;
        if (x > L) { x = 0.9999f * L; t = (vy > 0.f) ? (0.5f + dphi) * SBC_PI_F : -(0.5f + dphi) * SBC_PI_F; sincosf(t, &vy, &vx); hit = true; }
        else if (x < 0.f) { x = 0.0001f; t = (vy > 0.f) ? (0.5f - dphi) * SBC_PI_F : -(0.5f - dphi) * SBC_PI_F; sincosf(t, &vy, &vx); hit = true; }
        if (y > L) { y = 0.9999f * L; t = (vx > 0.f) ? -dphi : SBC_PI_F + dphi; sincosf(t, &vy, &vx); hit = true; }
        else if (y < 0.f) { y = 0.0001f; t = (vx > 0.f) ? dphi : SBC_PI_F - dphi; sincosf(t, &vy, &vx); hit = true; }
        return hit;
    }
    case 2: {  // elastic box :358-384
        bool hit = false;
        if (x > L) { x -= 2.f * (x - L); vx = -vx; hit = true; } else if (x < 0.f) { x -= 2.f * x; vx = -vx; hit = true; }
        if (y > L) { y -= 2.f * (y - L); vy = -vy; hit = true; } else if (y < 0.f) { y -= 2.f * y; vy = -vy; hit = true; }
        return hit;
    }
    case 3: {  // x periodic, y elastic :386-406 (the wrap counts as a boundary event, as in the oracle's flag)
        bool hit = false;
        if (x < 0.f) { x += L; hit = true; } else if (x >= L) { x -= L; hit = true; }
        if (y > L) { y -= 2.f * (y - L); vy = -vy; hit = true; } else if (y < 0.f) { y -= 2.f * y; vy = -vy; hit = true; }
        return hit;
    }
    case 4: {  // x periodic, y inelastic :407-430
        bool hit = false;
        if (x < 0.f) { x += L; hit = true; } else if (x >= L) { x -= L; hit = true; }
        if (y > L) { y = L - 0.0001f; vy = 0.f; hit = true; } else if (y < 0.f) { y = 0.0001f; vy = 0.f; hit = true; }
        return hit;
    }
    default:
        return false;
    }
}

// ---------------------------------------------------------------------------------------------
// ParticleBurstCoast, agents_dynamics.cpp:150-273, for one agent and one step.
//   acc       : accumulators of this row (only read when the agent starts a burst)
//   u_social, u_dir, k_next : this agent-step's decisions (Philox or injected)
// ---------------------------------------------------------------------------------------------
// (cos d, sin d) for |d| < 0.25 by Taylor polynomials (truncation error < 2e-11, i.e. below FP32
// rounding); used to turn the cached heading vector instead of calling sincosf every step
__device__ __forceinline__ void sbc_sincos_small(float d, float &s, float &c) {
    const float d2 = d * d;
    float ps = __fmaf_rn(d2, -1.f / 5040.f, 1.f / 120.f);
    ps = __fmaf_rn(d2, ps, -1.f / 6.f);
    s = __fmaf_rn(d * d2, ps, d);
    float pc = __fmaf_rn(d2, 1.f / 40320.f, -1.f / 720.f);
    pc = __fmaf_rn(d2, pc, 1.f / 24.f);
    pc = __fmaf_rn(d2, pc, -0.5f);
    c = __fmaf_rn(d2, pc, 1.f);
}

__device__ __forceinline__ void sbc_agent_step(AgentState &a, const RowAcc &acc, double u_social, float u_dir,
                                               uint32_t k_next, const DevParams &P, uint32_t &fl, const int BC) {
    float &cu = a.cu, &su = a.su;
    // (cu, su) = (cos, sin)(a.phi) on entry and on exit. The reference recomputes u from phi at
    // every step (:197,:227-228); here the cached vector is turned by the step's small heading
    // change (exact identity for a coasting agent, whose force is zero) and re-derived from phi
    // with sincosf at every burst start and at every rare event (wrap, wall hit, clamp), so the
    // drift is bounded by one burst (< 100 turns of ~1 ulp each).
    const bool first = (a.bin_step == P.burst_steps);
    const bool bursting = (a.bin_step > 0);
    float fx, fy, force_mag = P.soc;
    if (first) {
        fl |= SBC_FL_BURST_START;
        sincosf(a.phi, &su, &cu);
        sbc_decide(a, acc, u_social, u_dir, P, fx, fy, force_mag, fl, BC);
        a.fx = fx;
        a.fy = fy;
    } else if (bursting) {
        fx = a.fx;
        fy = a.fy;
    } else {
        fx = fy = 0.f;
        a.fx = a.fy = 0.f;
    }
    if (bursting) a.bin_step -= 1;
    if (a.stb > 0) a.stb -= 1;

    const float vproj_old = a.vproj;
    const float forcev = fx * cu + fy * su;
    float vproj = vproj_old + (-P.beta * vproj_old + forcev) * P.dt;  // :202
    float lphi = a.phi;
    float sl = su, cl = cu;
    const bool hack = vproj < 0.f;
    if (hack) {  // :206-210
        fl |= SBC_FL_VPROJ_HACK;
        vproj = 0.001f;
        lphi += 0.5f * SBC_PI_F;
        sincosf(lphi, &sl, &cl);
    }
    const float forcep = -fx * sl + fy * cl;
    const float turn = P.alphaTurn * forcep * P.dt * sbc_rcp_fast(vproj_old);  // :215 (old vproj)
    lphi += turn;
    float sn, cn;
    if (!hack && fabsf(turn) < 0.25f) {
        float sd, cd;
        sbc_sincos_small(turn, sd, cd);
        cn = __fmaf_rn(cu, cd, -su * sd);
        sn = __fmaf_rn(su, cd, cu * sd);
    } else {
        sincosf(lphi, &sn, &cn);
    }
    const float dphi = fabsf(lphi - a.phi);
    if (dphi > 0.01f) {
        // overshoot_check :123-137. acos is monotone decreasing, so the three angles are compared
        // through their cosines; an argument outside [-1,1] is NaN in the reference (every
        // comparison false) - reproduced by the `ok` predicates, with a rounding allowance of
        // 1e-6 so that |cos| = 1 + ulp is not mistaken for it.
        const float inv = (force_mag == P.soc) ? P.inv_soc : P.inv_env;
        const float c0 = (cu * fx + su * fy) * inv;  // cos(angle(u_old, F))
        const float c1 = (cn * fx + sn * fy) * inv;  // cos(angle(u_new, F))
        const float c01 = cn * cu + sn * su;         // cos(angle(u_new, u_old))
        const bool ok0 = fabsf(c0) <= 1.000001f, ok1 = fabsf(c1) <= 1.000001f;
        const bool o1 = ok0 && ok1 && (c0 > c1);
        const bool o2 = !o1 && ok0 && (c01 < c0);
        const bool o3 = dphi > SBC_PI_F;
        if (o1 || o2 || o3) {
            fl |= SBC_FL_OVERSHOOT;
            lphi = atan2f(fy, fx);
            sincosf(lphi, &sn, &cn);
        }
    }
    if (!(fabsf(lphi) < SBC_TWO_PI_F)) {  // :226 fmod keeps the sign; identity below 2 pi
        lphi = fmodf(lphi, SBC_TWO_PI_F);
        sincosf(lphi, &sn, &cn);
    }
    a.phi = lphi;
    cu = cn;
    su = sn;
    a.vproj = vproj;
    a.vx = vproj * cu;
    a.vy = vproj * su;
    a.x += a.vx * P.dt;
    a.y += a.vy * P.dt;
    bool hit = false;
    if (BC != -1) {  // :247
        hit = sbc_boundary(a.x, a.y, a.vx, a.vy, a.phi, P, BC);
        if (hit) fl |= SBC_FL_BOUNDARY;
    }
    if (a.stb == 0) {  // :250-270
        fl |= SBC_FL_REARM;
        a.bin_step = P.burst_steps;
        a.stb = k_next;
    }
    if (hit) {  // :272 - the second Boundary call can only act when the first one did
        sbc_boundary(a.x, a.y, a.vx, a.vy, a.phi, P, BC);
        sincosf(a.phi, &su, &cu);
    }
}
