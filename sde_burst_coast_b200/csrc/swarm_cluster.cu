// swarm_cluster.cu - mid-size single swarms (256 < N <= 2048) on a THREAD-BLOCK CLUSTER: one swarm is
// spread over up to 8 CTAs (256 agents each, thread = agent) that run on different SMs and read each
// other's agent positions through distributed shared memory (DSMEM). The whole Step()
// (/root/reference/src/swarmdyn.cpp:143-204) - interaction of the rows that start a burst, predator
// detection, burst decision, burst-coast update, boundary, predator spawn / move / kill - runs inside one
// persistent launch for many steps, with ONE cluster barrier per step, predators included. This is the
// Blackwell answer to swarms that are too large for one SM to step quickly (swarm_block.cu: ~13 us/step
// at N = 1024) and too small to amortise kernel launches (update.cu path).
//
// Data flow per step: every CTA publishes {x, y, vx, vy} of its agents in its own shared memory
// (double-buffered by step parity) -> cluster.sync() -> the warps of a CTA process the CTA's
// burst-starting rows against a local copy of the N partners (one bulk DSMEM read) -> thread-local update.
// The predator's move and kill of step s need the positions after that step's update, i.e. exactly what
// the exchange of step s + 1 publishes: they are evaluated there, by every CTA on its own copy of the
// whole swarm (centre of mass at spawn, closest prey, sensed / weight sums, dead set), so all CTAs hold
// identical predators and dead sets without a second barrier; a launch ends with one extra exchange
// for the predator phase of its last step.
#include <cooperative_groups.h>

#include "sbc_kernels.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int CT = 256;  // threads (= agents) per CTA

struct RowResult {
    float v0, v1, v2, v3;
    int cr, ca, ct, cf;
};
struct PredSchedule {
    uint32_t s_pred;
    int32_t s_trunc;
    int32_t needed;
};

__device__ __forceinline__ double warp_sum_d(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool PRED>
__global__ void __launch_bounds__(CT)
swarm_cluster_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps,
                     PredSchedule sched) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int N = S.N, NP = PRED ? S.Npred : 0;
    const int rep = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CT / 32;
    float4 *pv_s = reinterpret_cast<float4 *>(smem_raw);                  // [2][CT]
    RowResult *res_s = reinterpret_cast<RowResult *>(pv_s + 2 * CT);      // [CT]
    float2 *flee_s = reinterpret_cast<float2 *>(res_s + CT);              // [CT]
    double *wred_s = reinterpret_cast<double *>(flee_s + CT);             // [NW][5] warp partials
    unsigned long long *wkey_s = reinterpret_cast<unsigned long long *>(wred_s + NW * 5);  // [NW]
    unsigned *act_s = reinterpret_cast<unsigned *>(wkey_s + NW);          // [NW]
    int *n_rows_s = reinterpret_cast<int *>(act_s + NW);                  // [4]
    unsigned short *rows_s = reinterpret_cast<unsigned short *>(n_rows_s + 4);  // [CT]
    float *wsum_f = reinterpret_cast<float *>(rows_s + CT);               // [NW][8] warp partials of one row
    int *wsum_i = reinterpret_cast<int *>(wsum_f + NW * 8);                // [NW][4]
    float4 *all_s = reinterpret_cast<float4 *>(wsum_i + NW * 4);          // [C * CT] local copy of every CTA's positions
    float4 *pred_s = all_s + C * CT;                                      // [NP]
    float *pphi_s = reinterpret_cast<float *>(pred_s + NP);               // [NP]

    const int t = crank * CT + tid;   // agent index inside the swarm
    const bool valid = t < N;
    const size_t g = valid ? ((size_t)rep * N + t) : 0;
    AgentState a = {};
    bool alive = false, died = false;
    if (valid) {
        sbc_load_agent(P, S, g, a, true);
        alive = S.alive[g] != 0;
    }
    if (PRED) {
        for (int j = tid; j < NP; j += CT) {
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

    // sums of n <= 5 doubles per thread over this CTA (every thread gets the totals)
    auto cta_sum = [&](double *v, int n) {
        for (int k = 0; k < n; k++) {
            const double w = warp_sum_d(v[k]);
            if (lane == 0) wred_s[warp * 5 + k] = w;
        }
        __syncthreads();
        for (int k = 0; k < n; k++) {
            double tot = 0;
            for (int w = 0; w < NW; w++) tot += wred_s[w * 5 + k];
            v[k] = tot;
        }
        __syncthreads();
    };

    // CreatePredator / CreateFishNet (agents_dynamics.cpp:485-555) from the swarm copy in all_s; every CTA
    // computes the same predators
    auto spawn = [&](uint32_t step, uint32_t slot) {
        double v[5] = {0, 0, 0, 0, 0};
        const float scaling = SBC_TWO_PI_F * P.invL;
        for (int c = 0; c < C; c++) {
            const float4 p = all_s[c * CT + tid];
            if (p.x != p.x) continue;
            v[4] += 1.0;
            if (P.BC == 0) {
                float sn, cs;
                sincosf(scaling * p.x, &sn, &cs); v[0] += cs; v[1] += sn;
                sincosf(scaling * p.y, &sn, &cs); v[2] += cs; v[3] += sn;
            } else {
                v[0] += p.x;
                v[2] += p.y;
            }
        }
        cta_sum(v, 5);
        float comx, comy;
        if (P.BC == 0) {
            comx = (float)((atan2(-v[1] / v[4], -v[0] / v[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
            comy = (float)((atan2(-v[3] / v[4], -v[2] / v[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
        } else {
            comx = (float)(v[0] / v[4]);
            comy = (float)(v[2] / v[4]);
        }
        const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, step, slot);
        const float u0 = (float)((double)q.x * (1.0 / 4294967296.0));
        const float u1 = (float)((double)q.y * (1.0 / 4294967296.0));
        if (NP == 1) {
            if (tid == 0) {
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
            const float px = sn, py = -cs;
            const float net_len = 0.25f * P.L;
            sx -= 0.5f * net_len * px;
            sy -= 0.5f * net_len * py;
            const float net_dist = net_len / (float)(NP - 1);
            for (int j = tid; j < NP; j += CT) {
                float x = sx + px * ((float)j * net_dist), y = sy + py * ((float)j * net_dist);
                float vx = P.pred_speed * cs, vy = P.pred_speed * sn, ph = phi;
                sbc_boundary(x, y, vx, vy, ph, P, P.BC);
                pred_s[j] = make_float4(x, y, vx, vy);
                pphi_s[j] = ph;
            }
        }
        __syncthreads();
    };

    // Predator move + kill of step sp (swarmdyn.cpp:192-203), evaluated on the positions AFTER that step's
    // agent update - which are exactly what the next exchange publishes. Every CTA therefore runs this
    // phase on its own full copy of the swarm (all_s) right after the exchange: identical predators and
    // identical dead sets in every CTA without any further cluster barrier. A thread looks after the
    // agents c * CT + tid, c < C; the one with c == crank is its own.
    auto predator_phase = [&](uint32_t sp) {
        unsigned dead_mask = 0;   // bit c: agent c * CT + tid dies
        if (NP > 1) {
            for (int j = tid; j < NP; j += CT) {  // MoveFishNet :616-624
                float4 p = pred_s[j];
                float ph = pphi_s[j];
                p.x += p.z * P.dt;
                p.y += p.w * P.dt;
                sbc_boundary(p.x, p.y, p.z, p.w, ph, P, P.BC);
                pred_s[j] = p;
                pphi_s[j] = ph;
            }
            __syncthreads();
            if (P.pred_kill != 0) {  // FishNetKill :780-820
                const float4 p0 = pred_s[0], p1 = pred_s[1];
                const float ph0 = pphi_s[0];
                for (int c = 0; c < C; c++) {
                    const float4 p = all_s[c * CT + tid];
                    if (p.x == p.x && sbc_net_kills(p.x, p.y, p0, p1, ph0, NP, P)) dead_mask |= 1u << c;
                }
            }
        } else {
            const float4 p0 = pred_s[0];
            float lphi = pphi_s[0];
            if (P.pred_move == 0) {  // MovePredator :567-613, random walk
                const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, sp, 2u);
                const float u1 = (float)(((double)q.x + 1.0) * (1.0 / 4294967296.0));
                const float u2 = (float)((double)q.y * (1.0 / 4294967296.0));
                const float gss = sqrtf(-2.f * logf(u1)) * cosf(SBC_TWO_PI_F * u2);
                lphi = fmodf(lphi + P.noisep * gss * sqrtf(P.pred_speed), SBC_TWO_PI_F);
            } else {  // chase the closest prey: arg-min over the whole swarm inside the CTA
                unsigned long long best = ~0ull;
                for (int c = 0; c < C; c++) {
                    const float4 p = all_s[c * CT + tid];
                    if (p.x != p.x) continue;
                    float dx = p.x - p0.x, dy = p.y - p0.y;
                    if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, p.x, P); dy = sbc_delta_pbc(p0.y, p.y, P); }
                    best = min(best, ((unsigned long long)__float_as_uint(dx * dx + dy * dy) << 32) | (unsigned)(c * CT + tid));
                }
                for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
                if (lane == 0) wkey_s[warp] = best;
                __syncthreads();
                best = ~0ull;
                for (int w = 0; w < NW; w++) best = min(best, wkey_s[w]);
                if (best != ~0ull) {
                    const float4 p = all_s[(unsigned)(best & 0xffffffffu)];
                    float dx = p.x - p0.x, dy = p.y - p0.y;
                    if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, p.x, P); dy = sbc_delta_pbc(p0.y, p.y, P); }
                    lphi = atan2f(dy, dx);
                }
            }
            // every thread moves its own copy of the predator; thread 0 stores it
            float sn, cs;
            sincosf(lphi, &sn, &cs);
            float vx = P.pred_speed * cs, vy = P.pred_speed * sn;
            float x = p0.x + vx * P.dt, y = p0.y + vy * P.dt;
            float ph = lphi;
            sbc_boundary(x, y, vx, vy, ph, P, P.BC);
            __syncthreads();   // every thread has read pred_s[0]
            if (tid == 0) {
                pred_s[0] = make_float4(x, y, vx, vy);
                pphi_s[0] = ph;
            }
            if (P.pred_kill != 0) {  // PredKill :826-915
                unsigned cand_mask = 0;
                double v[2] = {0.0, 0.0};   // N_sensed, sum of the selection weights
                for (int c = 0; c < C; c++) {
                    const float4 p = all_s[c * CT + tid];
                    if (p.x != p.x) continue;
                    const KillClass kc = sbc_kill_classify(p.x, p.y, x, y, P);
                    if (P.pred_kill == 1) {
                        if (kc.within) dead_mask |= 1u << c;
                    } else if (kc.sensed) {
                        v[0] += 1.0;
                        if (kc.cand) {
                            cand_mask |= 1u << c;
                            if (P.pred_kill == 4) v[1] += sbc_kill_weight(p.x, p.y, x, y, P);
                        }
                    }
                }
                if (P.pred_kill > 1) {
                    cta_sum(v, 2);
                    for (int c = 0; c < C; c++) {
                        if (!((cand_mask >> c) & 1u)) continue;
                        const float4 p = all_s[c * CT + tid];
                        const double pk = sbc_kill_probability(p.x, p.y, x, y, (int)v[0], v[1], S.kill_eff, P);
                        const uint4 q = sbc_draw(P, rep, (uint32_t)(c * CT + tid), sp, SBC_SLOT_DECISION);
                        if (pk > (double)q.w * (1.0 / 4294967296.0)) dead_mask |= 1u << c;
                    }
                }
            }
        }
        for (int c = 0; c < C; c++)   // split_dead: the dead leave every CTA's copy of the swarm
            if ((dead_mask >> c) & 1u) all_s[c * CT + tid] = make_float4(qnan, qnan, 0.f, 0.f);
        if ((dead_mask >> crank) & 1u) {
            alive = false;
            died = true;
            n_killed++;
        }
        __syncthreads();
    };

    const uint32_t s_end = s_first + n_steps;
    for (uint32_t s = s_first;; s++) {
        // the predator phase of step s - 1 is still to do when this launch has run that step
        const bool pending = PRED && s != s_first && s != 0 && (s - 1 >= sched.s_pred);
        if (s == s_end && !pending) break;
        const bool pred_phase = PRED && (s >= sched.s_pred);
        const bool spawn0 = PRED && s != s_end && s == sched.s_pred;
        const bool spawn1 = PRED && s != s_end && NP > 1 && pred_phase && sched.needed > 0 &&
                            ((int)s - sched.s_trunc) % sched.needed == 0;
        float4 *my_pv = pv_s + (s & 1u) * CT;
        my_pv[tid] = alive ? make_float4(a.x, a.y, a.vx, a.vy) : make_float4(qnan, qnan, 0.f, 0.f);
        cluster.sync();   // the one cluster barrier of the step: everybody's positions are visible
        bool have_all = false;
        if (pending || spawn0 || spawn1) {
            for (int c = 0; c < C; c++)
                all_s[c * CT + tid] = (cluster.map_shared_rank(pv_s, c) + (s & 1u) * CT)[tid];
            __syncthreads();
            have_all = true;
            if (pending) predator_phase(s - 1);
            if (s == s_end) break;
            if (spawn0) { spawn(s, 0); spawned = true; }   // swarmdyn.cpp:153-158
            if (spawn1) spawn(s, 1);                        // :160-170
        }
        const bool first = alive && (a.bin_step == P.burst_steps);
        const unsigned m = __ballot_sync(FULL, first);
        if (lane == 0) act_s[warp] = m;
        __syncthreads();
        RowAcc acc;
        row_acc_clear(acc);
        // ordered list of this CTA's burst-starting rows
        if (warp == 0) {
            const unsigned mw = (lane < NW) ? act_s[lane] : 0u;
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
        if (n_first > 0 && !have_all) {
            // one bulk read of every CTA's positions through DSMEM (all loads of a thread in flight at once),
            // so that the row sweeps below run out of local shared memory
            for (int c = 0; c < C; c++)
                all_s[c * CT + tid] = (cluster.map_shared_rank(pv_s, c) + (s & 1u) * CT)[tid];
            __syncthreads();
        }
        // every burst-starting row is swept by ALL warps of the CTA (a warp takes every NW-th group of 32
        // partners), so the step's critical path is N / (32 NW) iterations instead of N / 32; the warp
        // partials are combined in warp order by one thread - a fixed order, independent of scheduling
        // predator detection: up to four predators share one Philox block and are handled by the row's own
        // thread after the sweeps (all rows in parallel); larger nets are spread over the lanes of a warp
        const bool row_detect = pred_phase && NP > 4;
        for (int kr = 0; kr < n_first; kr++) {
            const int row_l = rows_s[kr];             // index inside this CTA
            const int row = crank * CT + row_l;       // index inside the swarm
            const float4 pi = my_pv[row_l];
            int cr = 0, ca = 0, ct = 0;
            float rx = 0.f, ry = 0.f, gx = 0.f, gy = 0.f, tx = 0.f, ty = 0.f;
            for (int jl = tid; jl < N; jl += CT) {
                const float4 pj = all_s[jl];
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
                if (jl == row) z = 3;
                const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
                if (z == 0) { rx = __fmaf_rn(-dx, rinv, rx); ry = __fmaf_rn(-dy, rinv, ry); cr++; }
                else if (z == 1) { gx += pj.z; gy += pj.w; ca++; }
                else if (z == 2) { tx = __fmaf_rn(dx, rinv, tx); ty = __fmaf_rn(dy, rinv, ty); ct++; }
            }
            n_pairs += (tid == 0) ? (unsigned long long)(N - 1) : 0ull;
            cr = __reduce_add_sync(FULL, cr);
            ca = __reduce_add_sync(FULL, ca);
            ct = __reduce_add_sync(FULL, ct);
            float f0 = 0.f, f1 = 0.f;
            int cf = 0;
            if (row_detect && warp == (kr % NW)) {  // IntCalcPred (agents_interact.cpp:252-292): one warp per row
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
                rx += __shfl_xor_sync(FULL, rx, o); ry += __shfl_xor_sync(FULL, ry, o);
                gx += __shfl_xor_sync(FULL, gx, o); gy += __shfl_xor_sync(FULL, gy, o);
                tx += __shfl_xor_sync(FULL, tx, o); ty += __shfl_xor_sync(FULL, ty, o);
                if (PRED) {
                    f0 += __shfl_xor_sync(FULL, f0, o);
                    f1 += __shfl_xor_sync(FULL, f1, o);
                }
            }
            if (lane == 0) {
                float *wf = wsum_f + warp * 8;
                int *wi = wsum_i + warp * 4;
                wf[0] = rx; wf[1] = ry; wf[2] = gx; wf[3] = gy; wf[4] = tx; wf[5] = ty;
                wi[0] = cr; wi[1] = ca; wi[2] = ct;
                if (row_detect && warp == (kr % NW)) { flee_s[row_l] = make_float2(f0, f1); wi[3] = cf; }
            }
            __syncthreads();
            if (tid == 0) {
                float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                int c3[3] = {0, 0, 0};
                for (int w = 0; w < NW; w++) {
                    for (int q = 0; q < 6; q++) v[q] += wsum_f[w * 8 + q];
                    for (int q = 0; q < 3; q++) c3[q] += wsum_i[w * 4 + q];
                }
                RowResult r;
                r.v0 = (c3[0] > 0) ? v[0] : v[2]; r.v1 = (c3[0] > 0) ? v[1] : v[3]; r.v2 = v[4]; r.v3 = v[5];
                r.cr = c3[0]; r.ca = c3[1]; r.ct = c3[2];
                r.cf = row_detect ? wsum_i[(kr % NW) * 4 + 3] : 0;
                res_s[row_l] = r;
                if (!row_detect) flee_s[row_l] = make_float2(0.f, 0.f);
            }
            __syncthreads();
        }
        __syncthreads();
        if (first) {
            const RowResult r = res_s[tid];
            acc.c_rep = r.cr; acc.c_alg = r.ca; acc.c_att = r.ct; acc.c_flee = r.cf;
            if (r.cr > 0) { acc.rep_x = r.v0; acc.rep_y = r.v1; }
            else { acc.alg_x = r.v0; acc.alg_y = r.v1; }
            acc.att_x = r.v2; acc.att_y = r.v3;
            acc.flee_x = flee_s[tid].x; acc.flee_y = flee_s[tid].y;
            if (pred_phase && !row_detect) {  // IntCalcPred, agents_interact.cpp:252-292
                const uint4 q = sbc_draw(P, rep, (uint32_t)t, s, SBC_SLOT_DETECT0);
                const uint32_t r4[4] = {q.x, q.y, q.z, q.w};
                RowAcc fa;
                row_acc_clear(fa);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j < NP)
                        sbc_detect_accumulate(fa, a.x, a.y, pred_s[j].x, pred_s[j].y, (double)r4[j] * (1.0 / 4294967296.0), P);
                acc.c_flee = fa.c_flee;
                acc.flee_x = fa.flee_x;
                acc.flee_y = fa.flee_y;
            }
        }
        if (alive) {
            double u_social = 0.0;
            float u_dir = 0.f;
            uint32_t k_next = 1;
            const bool rearm = a.stb <= 1;
            if (first || rearm) {
                const uint4 q = sbc_draw(P, rep, (uint32_t)t, s, SBC_SLOT_DECISION);
                u_social = S.inj_social ? S.inj_social[g] : (double)q.x * (1.0 / 4294967296.0);
                u_dir = S.inj_dir ? S.inj_dir[g] : (float)(q.y >> 8) * (1.0f / 16777216.0f);
                if (rearm) k_next = S.inj_k ? S.inj_k[g] : sbc_geometric_k(S.geo, P.max_steps, q.z, P.inv_log2q);
            }
            uint32_t fl = 0;
            sbc_agent_step(a, acc, u_social, u_dir, k_next, P, fl, P.BC);
            flags_acc |= fl;
        }
    }
    if (valid && (alive || died)) {
        sbc_store_agent(P, S, g, a, true);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
        if (died) sbc_kill_agent(S, g, make_float4(a.x, a.y, a.vx, a.vy));
    }
    if (PRED) {
        __syncthreads();
        if (crank == 0) {
            for (int j = tid; j < NP; j += CT) {
                S.pred_pv[(size_t)rep * NP + j] = pred_s[j];
                S.pred_phi[(size_t)rep * NP + j] = pphi_s[j];
            }
            if (tid == 0 && spawned) S.pred_spawned[rep] = 1;
        }
        if (n_killed) atomicAdd(S.n_dead + rep, n_killed);
    }
    if (n_amb) atomicAdd(S.stats + 4, n_amb);
    if (n_pairs) atomicAdd(S.stats + 3, n_pairs);
    cluster.sync();   // no CTA exits while a neighbour may still read its shared memory
}

}  // namespace

bool sbc_cluster_supported(const DevParams &P, int N, int Npred) {
    return N > CT && N <= 8 * CT && Npred <= 1024 && !P.voronoi;
}

cudaError_t sbc_launch_swarm_cluster(const DevParams &P, const DevState &S, long long s_first, long long n_steps,
                                     cudaStream_t st, long long s_pred, int s_trunc, int needed) {
    const int C = (S.N + CT - 1) / CT;
    const size_t smem = 2 * CT * sizeof(float4) + CT * sizeof(RowResult) + CT * sizeof(float2) +
                        (CT / 32) * 5 * sizeof(double) + (CT / 32) * sizeof(unsigned long long) +
                        (CT / 32) * sizeof(unsigned) + 4 * sizeof(int) + CT * sizeof(unsigned short) +
                        (CT / 32) * 8 * sizeof(float) + (CT / 32) * 4 * sizeof(int) + (size_t)C * CT * sizeof(float4) + (size_t)S.Npred * (sizeof(float4) + sizeof(float)) + 64;
    PredSchedule sched;
    sched.s_pred = (s_pred < 0 || s_pred > 0xffffffffll) ? 0xffffffffu : (uint32_t)s_pred;
    sched.s_trunc = s_trunc;
    sched.needed = needed;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(S.R * C), 1, 1);
    cfg.blockDim = dim3(CT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint32_t s0 = (uint32_t)s_first, ns = (uint32_t)n_steps;
    if (S.Npred > 0) {
        cudaFuncSetAttribute(swarm_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return cudaLaunchKernelEx(&cfg, swarm_cluster_kernel<true>, P, S, s0, ns, sched);
    }
    cudaFuncSetAttribute(swarm_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaLaunchKernelEx(&cfg, swarm_cluster_kernel<false>, P, S, s0, ns, sched);
}
