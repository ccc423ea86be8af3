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
// when the predators of a swarm appear and re-appear, in step indices (host-evaluated, swarmdyn.cpp:153-170)
struct PredSchedule {
    uint32_t s_pred;    // first step with s >= pred_time/dt
    int32_t s_trunc;    // (int)(pred_time/dt), origin of the net's re-spawn period
    int32_t needed;     // re-spawn period in steps (fish net), 0 = never
};

constexpr int BC_RUNTIME = -99;   // template value: take the boundary condition from the parameters

// Work units of a launch. Classic: one warp per group of 32 / TPS tanks, all n_steps in one go - the last wave of warps
// leaves most SMs idle (16384 warps over 4736 resident ones: 3.46 waves run as 4). Sliced: the launch is cut into
// (group, time slice) units, one warp each, numbered slice-major so that the block scheduler itself hands them out in
// order; a unit's state goes through HBM (112 B per agent per slice: nothing next to a few hundred steps of arithmetic)
// and `done[group]` tells the slice that may run next - by the time a unit is dispatched, its group's previous slice
// was dispatched n_groups warps earlier and has normally long finished. With about a dozen units per resident warp the
// tail costs a per cent instead of an eighth. (The kernel body stays straight-line per warp: a persistent loop around
// it makes the compiler treat every shuffle as possibly divergent - 10 % slower.)
struct EnsSlices {
    int *done;           // [n_groups] slices of the group completed so far; nullptr: classic mode
    uint32_t slice;      // steps per unit
    unsigned n_groups;
};

template <int TPS, int BCT, bool VOR, bool SLICED>
__global__ void __launch_bounds__(128) __maxnreg__(VOR ? 128 : 64)
swarm_warp_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps, EnsSlices Q) {
    const int BC = (BCT == BC_RUNTIME) ? P.BC : BCT;
    constexpr int SPW = 32 / TPS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned unit = (unsigned)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    unsigned grp = unit, sl = 0;
    if (SLICED) {
        // the slice comes from the block index alone (a block's warps share it: n_groups is a multiple of the warps per
        // block), so that the step loop's bounds are uniform for the compiler - otherwise every shuffle inside it is
        // compiled as possibly divergent
        sl = (blockIdx.x * (blockDim.x >> 5)) / Q.n_groups;
        grp = unit - sl * Q.n_groups;
        s_first += sl * Q.slice;
        n_steps = (sl * Q.slice < n_steps) ? min(Q.slice, n_steps - sl * Q.slice) : 0u;
    }
    const int seg = lane / TPS, t = lane % TPS;
    const int rep = (int)grp * SPW + seg;
    const int N = S.N;
    const bool valid = (rep < S.R) && (t < N) && n_steps > 0;
    const size_t g = valid ? ((size_t)rep * N + t) : 0;
    const unsigned segmask = (TPS == 32) ? FULL : (((1u << TPS) - 1u) << (seg * TPS));
    AgentState a = {};
    bool alive = false;
    if (SLICED && sl > 0) {   // the group's previous slice
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(Q.done + grp) : "memory");
        } while (__any_sync(FULL, v < (int)sl));
    }
    if (valid) {
        if (SLICED) sbc_load_agent_cg(P, S, g, a);   // written by another SM during this launch: past L1
        else sbc_load_agent(P, S, g, a, true);
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
                sbc_delaunay_begin(e, xi, yi, sbc_vor_pos(xi, a.x, P), sbc_vor_pos(yi, a.y, P));
                const unsigned alive_mask = __ballot_sync(FULL, alive);
                const int seg0 = seg * TPS;
                for (int k = 0; k < TPS; k++) {
                    const int lk = seg0 + k;
                    const float xk = __shfl_sync(FULL, a.x, lk), yk = __shfl_sync(FULL, a.y, lk);
                    if (((alive_mask >> lk) & 1u) && lk != src && lk != lane) sbc_delaunay_add(e, sbc_vor_pos(xi, xk, P), sbc_vor_pos(yi, yk, P));
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
        sbc_store_agent(P, S, g, a, true);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
    }
    if (SLICED) {   // hand the group over to its next slice
        __threadfence();
        if (__all_sync(FULL, true) && lane == 0)   // (the vote doubles as the warp's execution barrier behind the fences)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(Q.done + grp), "r"((int)sl + 1) : "memory");
    }
    n_amb = __reduce_add_sync(FULL, (unsigned)n_amb);
    const unsigned np = __reduce_add_sync(FULL, (unsigned)n_pairs);
    if (lane == 0) {
        if (n_amb) atomicAdd(S.stats + 4, n_amb);
        if (np) atomicAdd(S.stats + 3, (unsigned long long)np);
    }
}

// The same with ONE predator per swarm (natPred, sinFisher: Npred == 1) carried in registers by every lane of the
// segment: spawn, detection by the burst-starting rows, chase / random walk, the four kill rules. Kept apart
// from the kernel above, whose code generation must not depend on it (it is the headline kernel of bench.py).
template <int TPS, int BCT, bool VOR>
__global__ void __launch_bounds__(128)
swarm_warp_pred_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps,
                  PredSchedule sched) {
    constexpr bool PRED = true;
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
    bool alive = false, died = false;
    if (valid) {
        sbc_load_agent(P, S, g, a, true);
        alive = S.alive[g] != 0;
    }
    uint32_t flags_acc = 0;
    unsigned long long n_amb = 0, n_pairs = 0;
    // PRED: the swarm's single predator (Npred == 1), replicated in every lane of the segment
    float4 pred = make_float4(0.f, 0.f, 0.f, 0.f);
    float pred_phi = 0.f;
    bool spawned = false, phi_stale = false;
    if (PRED && rep < S.R) {
        pred = S.pred_pv[rep];
        pred_phi = S.pred_phi[rep];
    }

    for (uint32_t s = s_first; s != s_first + n_steps; s++) {
        const bool pred_phase = PRED && (s >= sched.s_pred);
        if (PRED && s == sched.s_pred) {  // CreatePredator (agents_dynamics.cpp:534-555, swarmdyn.cpp:153-158)
            double c[5] = {0, 0, 0, 0, alive ? 1.0 : 0.0};
            if (alive) {
                if (BC == 0) {
                    const float scaling = SBC_TWO_PI_F * P.invL;
                    float sn, cs;
                    sincosf(scaling * a.x, &sn, &cs); c[0] = cs; c[1] = sn;
                    sincosf(scaling * a.y, &sn, &cs); c[2] = cs; c[3] = sn;
                } else {
                    c[0] = a.x;
                    c[2] = a.y;
                }
            }
#pragma unroll
            for (int k = 0; k < 5; k++)
                for (int o = TPS / 2; o > 0; o >>= 1) c[k] += __shfl_xor_sync(FULL, c[k], o);
            float comx, comy;
            if (BC == 0) {
                comx = (float)((atan2(-c[1] / c[4], -c[0] / c[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
                comy = (float)((atan2(-c[3] / c[4], -c[2] / c[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
            } else {
                comx = (float)(c[0] / c[4]);
                comy = (float)(c[2] / c[4]);
            }
            const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, s, 0u);
            float phi = (float)((double)q.x * (1.0 / 4294967296.0)) * SBC_TWO_PI_F, sn, cs;
            sincosf(phi, &sn, &cs);
            pred.x = comx - 0.5f * P.L * cs;
            pred.y = comy - 0.5f * P.L * sn;
            pred.z = P.pred_speed * cs;
            pred.w = P.pred_speed * sn;
            sbc_boundary(pred.x, pred.y, pred.z, pred.w, phi, P, BC);
            pred_phi = phi;
            spawned = true;
        }
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
                sbc_delaunay_begin(e, xi, yi, sbc_vor_pos(xi, a.x, P), sbc_vor_pos(yi, a.y, P));
                const unsigned alive_mask = __ballot_sync(FULL, alive);
                const int seg0 = seg * TPS;
                for (int k = 0; k < TPS; k++) {
                    const int lk = seg0 + k;
                    const float xk = __shfl_sync(FULL, a.x, lk), yk = __shfl_sync(FULL, a.y, lk);
                    if (((alive_mask >> lk) & 1u) && lk != src && lk != lane) sbc_delaunay_add(e, sbc_vor_pos(xi, xk, P), sbc_vor_pos(yi, yk, P));
                }
                if (pred_phase) sbc_delaunay_add(e, sbc_vor_pos(xi, pred.x, P), sbc_vor_pos(yi, pred.y, P));   // a vertex of F2FP's triangulation
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
        if (PRED && pred_phase && first) {  // IntCalcPred (agents_interact.cpp:252-292): the row's own draw
            const uint4 q = sbc_draw(P, rep, (uint32_t)t, s, SBC_SLOT_DETECT0);
            RowAcc fa;
            row_acc_clear(fa);
            sbc_detect_accumulate(fa, a.x, a.y, pred.x, pred.y, (double)q.x * (1.0 / 4294967296.0), P);
            acc.c_flee = fa.c_flee;
            acc.flee_x = fa.flee_x;
            acc.flee_y = fa.flee_y;
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
        if (PRED && pred_phase) {  // swarmdyn.cpp:192-203, every lane of the segment moves its copy of the predator
            float lphi = pred_phi;
            bool chase_hit = false;
            float chase_cs = 1.f, chase_sn = 0.f;
            if (P.pred_move == 0) {  // MovePredator :567-613, random walk
                const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, s, 2u);
                const float u1 = (float)(((double)q.x + 1.0) * (1.0 / 4294967296.0));
                const float u2 = (float)((double)q.y * (1.0 / 4294967296.0));
                const float gss = sqrtf(-2.f * logf(u1)) * cosf(SBC_TWO_PI_F * u2);
                lphi = fmodf(pred_phi + P.noisep * gss * sqrtf(P.pred_speed), SBC_TWO_PI_F);
            } else {  // chase the closest prey: arg-min of (d2, index) over the segment
                unsigned key = 0xffffffffu;   // d2 >= 0: its bit pattern orders like the value
                float dx = 0.f, dy = 0.f;
                if (alive) {
                    dx = a.x - pred.x; dy = a.y - pred.y;
                    if (BC == 0) { dx = sbc_delta_pbc(pred.x, a.x, P); dy = sbc_delta_pbc(pred.y, a.y, P); }
                    key = __float_as_uint(dx * dx + dy * dy);
                }
                const unsigned best = __reduce_min_sync(segmask, key);
                const unsigned winners = __ballot_sync(FULL, key == best) & segmask;   // lowest index wins a tie
                const int wl = __ffs(winners) - 1;
                const float wdx = __shfl_sync(FULL, dx, wl), wdy = __shfl_sync(FULL, dy, wl);
                chase_hit = best != 0xffffffffu;
                if (BCT == 0) {
                    // periodic box: the heading never changes at the boundary, so the velocity
                    // speed * (cos, sin)(atan2(dy, dx)) is taken as speed * (dx, dy) / |d| and the angle
                    // itself is only formed when the launch ends
                    const float d2w = wdx * wdx + wdy * wdy;
                    const float rinv = rsqrtf(d2w);
                    chase_cs = (d2w > 0.f) ? wdx * rinv : 1.f;
                    chase_sn = (d2w > 0.f) ? wdy * rinv : 0.f;
                } else if (chase_hit) {
                    lphi = atan2f(wdy, wdx);
                }
            }
            if (BCT == 0 && P.pred_move != 0) {
                if (chase_hit) {
                    pred.z = P.pred_speed * chase_cs;
                    pred.w = P.pred_speed * chase_sn;
                    phi_stale = true;
                }
                pred.x += pred.z * P.dt;
                pred.y += pred.w * P.dt;
                float ph = 0.f;
                sbc_boundary(pred.x, pred.y, pred.z, pred.w, ph, P, BC);
            } else {
                float sn, cs;
                sincosf(lphi, &sn, &cs);
                pred.z = P.pred_speed * cs;
                pred.w = P.pred_speed * sn;
                pred.x += pred.z * P.dt;
                pred.y += pred.w * P.dt;
                pred_phi = lphi;
                sbc_boundary(pred.x, pred.y, pred.z, pred.w, pred_phi, P, BC);
            }
            if (P.pred_kill != 0) {  // PredKill :826-915
                bool die = false;
                KillClass kc = {false, false, false};
                if (alive) kc = sbc_kill_classify(a.x, a.y, pred.x, pred.y, P);
                if (P.pred_kill == 1) {
                    die = kc.within;
                } else {
                    const int n_sensed = __popc(__ballot_sync(FULL, kc.sensed) & segmask);
                    double total = 0.0;
                    if (P.pred_kill == 4) {
                        total = kc.cand ? sbc_kill_weight(a.x, a.y, pred.x, pred.y, P) : 0.0;
#pragma unroll
                        for (int o = TPS / 2; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
                    }
                    if (kc.cand) {
                        const double pk = sbc_kill_probability(a.x, a.y, pred.x, pred.y, n_sensed, total, S.kill_eff, P);
                        const uint4 q = sbc_draw(P, rep, (uint32_t)t, s, SBC_SLOT_DECISION);
                        die = pk > (double)q.w * (1.0 / 4294967296.0);
                    }
                }
                if (die) {
                    alive = false;
                    died = true;
                }
            }
        }
    }
    if (valid && (alive || died)) {
        sbc_store_agent(P, S, g, a, true);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
        if (died) {  // split_dead (agents_operation.cpp:209-218)
            sbc_kill_agent(S, g, make_float4(a.x, a.y, a.vx, a.vy));
            atomicAdd(S.n_dead + rep, 1);
        }
    }
    if (PRED && rep < S.R && t == 0) {
        if (phi_stale) pred_phi = atan2f(pred.w, pred.z);
        S.pred_pv[rep] = pred;
        S.pred_phi[rep] = pred_phi;
        if (spawned) S.pred_spawned[rep] = 1;
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
    int cr, ca, ct, cf;
};

// block-wide helpers (nwarps <= 32); every thread of the block must call them
__device__ __forceinline__ double block_sum_d(double v, double *sh, int nwarps) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < nwarps; w++) t += sh[w];
    return t;
}
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, unsigned long long *sh, int nwarps) {
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = ~0ull;
    for (int w = 0; w < nwarps; w++) t = min(t, sh[w]);
    return t;
}

// CreatePredator / CreateFishNet (agents_dynamics.cpp:534-555, 485-528) by the whole block:
// GetCenterOfMass (agents_operation.cpp:23-127) of the alive agents, then the predators into pred_s
__device__ void block_spawn(const DevParams &P, int NP, uint32_t rep, uint32_t step, uint32_t slot, bool alive,
                            const AgentState &a, float4 *pred_s, float *pphi_s, double *red_s, int nwarps) {
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, cnt = alive ? 1.0 : 0.0;
    if (alive) {
        if (P.BC == 0) {
            const float scaling = SBC_TWO_PI_F * P.invL;
            float sn, cs;
            sincosf(scaling * a.x, &sn, &cs); c0 = cs; c1 = sn;
            sincosf(scaling * a.y, &sn, &cs); c2 = cs; c3 = sn;
        } else {
            c0 = a.x;
            c2 = a.y;
        }
    }
    c0 = block_sum_d(c0, red_s, nwarps); c1 = block_sum_d(c1, red_s, nwarps);
    c2 = block_sum_d(c2, red_s, nwarps); c3 = block_sum_d(c3, red_s, nwarps);
    cnt = block_sum_d(cnt, red_s, nwarps);
    float comx, comy;
    if (P.BC == 0) {
        comx = (float)((atan2(-c1 / cnt, -c0 / cnt) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
        comy = (float)((atan2(-c3 / cnt, -c2 / cnt) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
    } else {
        comx = (float)(c0 / cnt);
        comy = (float)(c2 / cnt);
    }
    const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, step, slot);
    const float u0 = (float)((double)q.x * (1.0 / 4294967296.0));
    const float u1 = (float)((double)q.y * (1.0 / 4294967296.0));
    if (NP == 1) {
        if (threadIdx.x == 0) {
            float phi = u0 * SBC_TWO_PI_F, sn, cs;
            sincosf(phi, &sn, &cs);
            float x = comx - 0.5f * P.L * cs, y = comy - 0.5f * P.L * sn;
            float vx = P.pred_speed * cs, vy = P.pred_speed * sn;
            sbc_boundary(x, y, vx, vy, phi, P, P.BC);
            pred_s[0] = make_float4(x, y, vx, vy);
            pphi_s[0] = phi;
        }
    } else {
        float phi = SBC_TWO_PI_F * u0, sn, cs;
        sincosf(phi, &sn, &cs);
        float sx = comx - 0.25f * P.L * cs, sy = comy - 0.25f * P.L * sn;
        const float var_noise = SBC_PI_F / 4.f;
        phi += var_noise * u1 - 0.5f * var_noise;
        sincosf(phi, &sn, &cs);
        const float px = sn, py = -cs;  // vec_perp(v_net_move), mathtools.h:183-189
        const float net_len = 0.25f * P.L;
        sx -= 0.5f * net_len * px;
        sy -= 0.5f * net_len * py;
        const float net_dist = net_len / (float)(NP - 1);
        for (int j = threadIdx.x; j < NP; j += blockDim.x) {
            float x = sx + px * ((float)j * net_dist), y = sy + py * ((float)j * net_dist);
            float vx = P.pred_speed * cs, vy = P.pred_speed * sn, ph = phi;
            sbc_boundary(x, y, vx, vy, ph, P, P.BC);
            pred_s[j] = make_float4(x, y, vx, vy);
            pphi_s[j] = ph;
        }
    }
    __syncthreads();
}

template <bool PRED, int MAXT>
__global__ void __launch_bounds__(MAXT)
swarm_block_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps,
                   PredSchedule sched) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    const int N = S.N, NP = PRED ? S.Npred : 0;
    const int nthreads = blockDim.x;
    const int nwarps = nthreads >> 5;
    float4 *pv_s = reinterpret_cast<float4 *>(smem_raw);                       // [nthreads]
    RowResult *res_s = reinterpret_cast<RowResult *>(pv_s + nthreads);         // [nthreads]
    float2 *flee_s = reinterpret_cast<float2 *>(res_s + nthreads);             // [nthreads]
    double *red_s = reinterpret_cast<double *>(flee_s + nthreads);             // [32]
    unsigned long long *key_s = reinterpret_cast<unsigned long long *>(red_s + 32);  // [32]
    unsigned *act_s = reinterpret_cast<unsigned *>(key_s + 32);                // [32]
    int *n_rows_s = reinterpret_cast<int *>(act_s + 32);                       // [4]
    unsigned short *rows_s = reinterpret_cast<unsigned short *>(n_rows_s + 4); // [nthreads]
    float4 *pred_s = reinterpret_cast<float4 *>(rows_s + ((nthreads + 7) / 8) * 8);  // [NP], 16-byte aligned
    float *pphi_s = reinterpret_cast<float *>(pred_s + NP);                    // [NP]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int rep = blockIdx.x;
    const bool valid = t < N;
    const size_t g = valid ? ((size_t)rep * N + t) : 0;
    AgentState a = {};
    bool alive = false, died = false;
    if (valid) {
        sbc_load_agent(P, S, g, a, true);
        alive = S.alive[g] != 0;
    }
    if (PRED) {
        for (int j = t; j < NP; j += nthreads) {
            pred_s[j] = S.pred_pv[(size_t)rep * NP + j];
            pphi_s[j] = S.pred_phi[(size_t)rep * NP + j];
        }
        __syncthreads();
    }
    uint32_t flags_acc = 0;
    unsigned long long n_amb = 0, n_pairs = 0;
    int n_killed = 0;
    bool spawned = false;
    const float qnan = __int_as_float(0x7fc00000);

    for (uint32_t s = s_first; s != s_first + n_steps; s++) {
        const bool pred_phase = PRED && (s >= sched.s_pred);
        if (PRED) {
            if (s == sched.s_pred) {  // swarmdyn.cpp:153-158
                block_spawn(P, NP, rep, s, 0, alive, a, pred_s, pphi_s, red_s, nwarps);
                spawned = true;
            }
            if (NP > 1 && pred_phase && sched.needed > 0 && ((int)s - sched.s_trunc) % sched.needed == 0)
                block_spawn(P, NP, rep, s, 1, alive, a, pred_s, pphi_s, red_s, nwarps);  // :160-170
        }
        const bool first = alive && (a.bin_step == P.burst_steps);
        RowAcc acc;
        row_acc_clear(acc);
        const int any = __syncthreads_or(first ? 1 : 0);
        if (any) {
            pv_s[t] = alive ? make_float4(a.x, a.y, a.vx, a.vy) : make_float4(qnan, qnan, 0.f, 0.f);
            const unsigned m = __ballot_sync(FULL, first);
            if (lane == 0) act_s[warp] = m;
            __syncthreads();
            // warp 0 turns the per-warp masks into an ordered list of the rows that start a burst
            if (warp == 0) {
                const unsigned mw = (lane < nwarps) ? act_s[lane] : 0u;
                const int c = __popc(mw);
                int off = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, off, o);
                    if (lane >= o) off += v;
                }
                if (lane == 31) n_rows_s[0] = off;
                off -= c;
                unsigned m2 = mw;
                while (m2) {
                    const int bit = __ffs(m2) - 1;
                    m2 &= m2 - 1;
                    rows_s[off++] = (unsigned short)(lane * 32 + bit);
                }
            }
            __syncthreads();
            const int n_first = n_rows_s[0];
            {
                for (int kr = warp; kr < n_first; kr += nwarps) {
                    const int row = rows_s[kr];
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
                            sbc_delaunay_begin(e, pi.x, pi.y, sbc_vor_pos(pi.x, pj.x, P), sbc_vor_pos(pi.y, pj.y, P));
                            for (int kk = 0; kk < N; kk++) {
                                const float4 pk = pv_s[kk];
                                if (kk != row && kk != j && pk.x == pk.x) sbc_delaunay_add(e, sbc_vor_pos(pi.x, pk.x, P), sbc_vor_pos(pi.y, pk.y, P));
                            }
                            if (pred_phase)   // the predators are vertices of F2FP's triangulation (agents_interact.cpp:103-110)
                                for (int q = 0; q < NP; q++) {
                                    const float4 pq = S.pred_pv[(size_t)rep * NP + q];
                                    sbc_delaunay_add(e, sbc_vor_pos(pi.x, pq.x, P), sbc_vor_pos(pi.y, pq.y, P));
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
                    float f0 = 0.f, f1 = 0.f;
                    int cf = 0;
                    if (pred_phase) {  // IntCalcPred for this row: a lane takes four predators per Philox block
                        RowAcc fa;
                        row_acc_clear(fa);
                        for (int jb = lane; 4 * jb < NP; jb += 32) {
                            const uint4 q = sbc_draw(P, rep, (uint32_t)row, s, SBC_SLOT_DETECT0 + (uint32_t)jb);
                            const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int j = 4 * jb + u;
                                if (j < NP)
                                    sbc_detect_accumulate(fa, pi.x, pi.y, pred_s[j].x, pred_s[j].y,
                                                          (double)r4[u] * (1.0 / 4294967296.0), P);
                            }
                        }
                        f0 = fa.flee_x; f1 = fa.flee_y;
                        cf = __reduce_add_sync(FULL, fa.c_flee);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        v0 += __shfl_xor_sync(FULL, v0, o);
                        v1 += __shfl_xor_sync(FULL, v1, o);
                        v2 += __shfl_xor_sync(FULL, v2, o);
                        v3 += __shfl_xor_sync(FULL, v3, o);
                        if (PRED) {
                            f0 += __shfl_xor_sync(FULL, f0, o);
                            f1 += __shfl_xor_sync(FULL, f1, o);
                        }
                    }
                    if (lane == 0) {
                        RowResult r;
                        r.v0 = v0; r.v1 = v1; r.v2 = v2; r.v3 = v3;
                        r.cr = cr; r.ca = ca; r.ct = ct; r.cf = cf;
                        res_s[row] = r;
                        flee_s[row] = make_float2(f0, f1);
                    }
                }
            }
            __syncthreads();
            if (first) {
                const RowResult r = res_s[t];
                acc.c_rep = r.cr; acc.c_alg = r.ca; acc.c_att = r.ct; acc.c_flee = r.cf;
                if (r.cr > 0) { acc.rep_x = r.v0; acc.rep_y = r.v1; }
                else { acc.alg_x = r.v0; acc.alg_y = r.v1; }
                acc.att_x = r.v2; acc.att_y = r.v3;
                acc.flee_x = flee_s[t].x; acc.flee_y = flee_s[t].y;
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
        if (pred_phase) {  // swarmdyn.cpp:192-203
            bool die = false;
            if (NP > 1) {
                __syncthreads();
                for (int j = t; j < NP; j += nthreads) {  // MoveFishNet :616-624
                    float4 p = pred_s[j];
                    float ph = pphi_s[j];
                    p.x += p.z * P.dt;
                    p.y += p.w * P.dt;
                    sbc_boundary(p.x, p.y, p.z, p.w, ph, P, P.BC);
                    pred_s[j] = p;
                    pphi_s[j] = ph;
                }
                __syncthreads();
                if (P.pred_kill != 0 && alive)  // FishNetKill :780-820
                    die = sbc_net_kills(a.x, a.y, pred_s[0], pred_s[1], pphi_s[0], NP, P);
            } else {
                // MovePredator :567-613
                const float4 p0 = pred_s[0];
                float lphi = pphi_s[0];
                if (P.pred_move == 0) {
                    const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, s, 2u);
                    const float u1 = (float)(((double)q.x + 1.0) * (1.0 / 4294967296.0));
                    const float u2 = (float)((double)q.y * (1.0 / 4294967296.0));
                    const float gss = sqrtf(-2.f * logf(u1)) * cosf(SBC_TWO_PI_F * u2);
                    lphi = fmodf(pphi_s[0] + P.noisep * gss * sqrtf(P.pred_speed), SBC_TWO_PI_F);
                } else {
                    unsigned long long key = ~0ull;
                    float dx = 0.f, dy = 0.f;
                    if (alive) {
                        dx = a.x - p0.x; dy = a.y - p0.y;
                        if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, a.x, P); dy = sbc_delta_pbc(p0.y, a.y, P); }
                        key = ((unsigned long long)__float_as_uint(dx * dx + dy * dy) << 32) | (unsigned)t;
                    }
                    const unsigned long long best = block_min_u64(key, key_s, nwarps);
                    __syncthreads();
                    if (best != ~0ull && (unsigned)(best & 0xffffffffu) == (unsigned)t) red_s[0] = (double)atan2f(dy, dx);
                    __syncthreads();
                    if (best != ~0ull) lphi = (float)red_s[0];
                }
                __syncthreads();
                if (t == 0) {
                    float sn, cs;
                    sincosf(lphi, &sn, &cs);
                    float vx = P.pred_speed * cs, vy = P.pred_speed * sn;
                    float x = p0.x + vx * P.dt, y = p0.y + vy * P.dt;
                    float ph = lphi;
                    sbc_boundary(x, y, vx, vy, ph, P, P.BC);
                    pred_s[0] = make_float4(x, y, vx, vy);
                    pphi_s[0] = ph;
                }
                __syncthreads();
                if (P.pred_kill != 0) {  // PredKill :826-915
                    const float4 pp = pred_s[0];
                    KillClass kc = {false, false, false};
                    if (alive) kc = sbc_kill_classify(a.x, a.y, pp.x, pp.y, P);
                    if (P.pred_kill == 1) {
                        die = kc.within;
                    } else {
                        const int n_sensed = __syncthreads_count(kc.sensed ? 1 : 0);
                        double total = 0.0;
                        if (P.pred_kill == 4)
                            total = block_sum_d(kc.cand ? sbc_kill_weight(a.x, a.y, pp.x, pp.y, P) : 0.0, red_s, nwarps);
                        if (kc.cand) {
                            const double pk = sbc_kill_probability(a.x, a.y, pp.x, pp.y, n_sensed, total, S.kill_eff, P);
                            const uint4 q = sbc_draw(P, rep, (uint32_t)t, s, SBC_SLOT_DECISION);
                            die = pk > (double)q.w * (1.0 / 4294967296.0);
                        }
                    }
                }
            }
            if (die) {
                alive = false;
                died = true;
                n_killed++;
            }
        }
    }
    if (valid && (alive || died)) {
        sbc_store_agent(P, S, g, a, true);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
        if (died) sbc_kill_agent(S, g, make_float4(a.x, a.y, a.vx, a.vy));
    }
    if (PRED) {
        __syncthreads();
        for (int j = t; j < NP; j += nthreads) {
            S.pred_pv[(size_t)rep * NP + j] = pred_s[j];
            S.pred_phi[(size_t)rep * NP + j] = pphi_s[j];
        }
        if (n_killed) atomicAdd(S.n_dead + rep, n_killed);
        if (t == 0 && spawned) S.pred_spawned[rep] = 1;
    }
    if (n_amb) atomicAdd(S.stats + 4, n_amb);
    if (n_pairs) atomicAdd(S.stats + 3, n_pairs);
}

}  // namespace

bool sbc_launch_swarm_block(const DevParams &P, const DevState &S, long long s_first, long long n_steps,
                            cudaStream_t st, long long s_pred, int s_trunc, int needed, bool force_block) {
    const int N = S.N;
    if (N > 1024 || N < 1 || S.Npred > 1024) return false;
    const uint32_t s0 = (uint32_t)s_first, ns = (uint32_t)n_steps;
    PredSchedule sched;
    sched.s_pred = (s_pred < 0 || s_pred > 0xffffffffll) ? 0xffffffffu : (uint32_t)s_pred;
    sched.s_trunc = s_trunc;
    sched.needed = needed;
    if (N <= 32 && S.Npred <= 1 && !(force_block && S.Npred == 1)) {
        const int tps = (N <= 8) ? 8 : (N <= 16) ? 16 : 32;
        const int spw = 32 / tps;
        const long long warps = ((long long)S.R + spw - 1) / spw;
        const int wpb = 4;
        const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
        // sliced mode (see EnsSlices) when the launch is a few waves of long-running warps: units of a time slice each,
        // about a dozen per resident warp
        EnsSlices Q = {nullptr, ns, (unsigned)warps};
        unsigned qblocks = blocks;
        {
            static int sms = 0;
            if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
            const long long resident = (long long)sms * 8 * wpb;   // 8 blocks of 4 warps per SM at 64 registers
            const char *env = getenv("SBC_ENS_QUEUE");
            const bool allow = !(env && env[0] == '0') && S.Npred == 0 && S.ens_done != nullptr && warps % wpb == 0 && !P.voronoi;
            long long n_slices = (12 * resident + warps - 1) / warps;
            if (const char *e2 = getenv("SBC_ENS_SLICES")) n_slices = atoll(e2);   // tuning runs
            if (n_slices > (long long)ns / 64) n_slices = (long long)ns / 64;
            if (allow && warps > resident && n_slices > 1) {
                Q.done = S.ens_done;
                Q.slice = (uint32_t)((ns + n_slices - 1) / n_slices);
                n_slices = (ns + Q.slice - 1) / Q.slice;
                qblocks = (unsigned)(warps * n_slices / wpb);
                cudaMemsetAsync(S.ens_done, 0, (size_t)warps * sizeof(int), st);
            }
        }
#define SBC_WARP_LAUNCH(T, B, PR)                                                                         \
    do {                                                                                                  \
        if (PR) {                                                                                         \
            if (P.voronoi) swarm_warp_pred_kernel<T, B, true><<<blocks, wpb * 32, 0, st>>>(P, S, s0, ns, sched); \
            else swarm_warp_pred_kernel<T, B, false><<<blocks, wpb * 32, 0, st>>>(P, S, s0, ns, sched);  \
        } else {                                                                                          \
            if (P.voronoi) swarm_warp_kernel<T, B, true, false><<<blocks, wpb * 32, 0, st>>>(P, S, s0, ns, Q);     \
            else if (Q.done) swarm_warp_kernel<T, B, false, true><<<qblocks, wpb * 32, 0, st>>>(P, S, s0, ns, Q);  \
            else swarm_warp_kernel<T, B, false, false><<<blocks, wpb * 32, 0, st>>>(P, S, s0, ns, Q);              \
        }                                                                                                 \
    } while (0)
#define SBC_WARP_BC(T)                                                                     \
    do {                                                                                   \
        if (S.Npred == 1) {  /* predator variants: periodic box or run-time boundary */    \
            if (P.BC == 0) SBC_WARP_LAUNCH(T, 0, true);                                    \
            else SBC_WARP_LAUNCH(T, BC_RUNTIME, true);                                     \
        } else if (P.BC == 6) SBC_WARP_LAUNCH(T, 6, false);                                \
        else if (P.BC == 0) SBC_WARP_LAUNCH(T, 0, false);                                  \
        else SBC_WARP_LAUNCH(T, BC_RUNTIME, false);                                        \
    } while (0)
        if (tps == 8) SBC_WARP_BC(8);
        else if (tps == 16) SBC_WARP_BC(16);
        else SBC_WARP_BC(32);
#undef SBC_WARP_BC
#undef SBC_WARP_LAUNCH
        return true;
    }
    const int nthreads = ((N + 31) / 32) * 32;
    const size_t smem = (size_t)nthreads * (sizeof(float4) + sizeof(RowResult) + sizeof(float2)) + 32 * sizeof(double) +
                        32 * sizeof(unsigned long long) + 32 * sizeof(unsigned) + 4 * sizeof(int) + (size_t)((nthreads + 7) / 8) * 8 * sizeof(unsigned short) + (size_t)S.Npred * (sizeof(float4) + sizeof(float)) + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(swarm_block_kernel<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(swarm_block_kernel<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(swarm_block_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(swarm_block_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set = true;
    }
    // two register budgets: swarms of up to 256 agents get the larger one
    if (nthreads <= 256) {
        if (S.Npred > 0) swarm_block_kernel<true, 256><<<S.R, nthreads, smem, st>>>(P, S, s0, ns, sched);
        else swarm_block_kernel<false, 256><<<S.R, nthreads, smem, st>>>(P, S, s0, ns, sched);
    } else {
        if (S.Npred > 0) swarm_block_kernel<true, 1024><<<S.R, nthreads, smem, st>>>(P, S, s0, ns, sched);
        else swarm_block_kernel<false, 1024><<<S.R, nthreads, smem, st>>>(P, S, s0, ns, sched);
    }
    return true;
}
