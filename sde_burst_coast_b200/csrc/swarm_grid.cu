// swarm_grid.cu - single swarms of a few thousand to a few ten thousand agents (2048 < N <= 256 x the number of
// co-resident CTAs, 37,888 on a B200) in ONE persistent cooperative launch: BASELINE config C4 (mulFisher,
// N = 16,384). The multi-kernel path spends ~4 us per kernel boundary and needs seven kernels per step there;
// here a step is one grid barrier.
//
// Thread = agent, state in registers for the whole launch (as in swarm_block.cu / swarm_cluster.cu). Per step:
//   every thread publishes {x, y, vx, vy} of its agent in a global position buffer (two buffers by step
//   parity, 16 B per agent: the whole swarm stays in L2) -> grid.sync() -> the CTA's burst-starting rows are
//   dealt to its warps and swept against all N published positions, which stream through shared memory in
//   512-agent tiles (cp.async.cg, four stages in flight; a cheap FP32 far test rejects most pairs before the
//   exact periodic difference is formed), prey<-net detection for those rows (IntCalcPred,
//   /root/reference/src/agents_interact.cpp:252-292) -> burst decision + burst-coast update + boundary in
//   registers (sbc_agent_step) -> fish net: every CTA moves its own copy of the net nodes (MoveFishNet,
//   agents_dynamics.cpp:616-624) and each thread tests its own agent (FishNetKill :780-820).
// Positions written by other SMs are read with cp.async.cg / ld.global.cg (L2): L1 is not coherent across SMs
// and the same addresses are rewritten every second step.
// Not handled here (sbc_grid_supported says so, the multi-kernel path takes over): a single chasing predator
// (needs swarm-wide arg-min and sums after the update), Voronoi neighbour rule, ensembles.
#include <cooperative_groups.h>

#include "sbc_kernels.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int GT = 256;  // threads (= agents) per CTA
constexpr int NW = GT / 32;
constexpr int TILE = 512;    // agents per shared-memory tile of the sweep
constexpr int STAGES = 4;    // tiles in flight
constexpr int RPW = 4;       // rows a warp sweeps in one pass over the swarm

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
__global__ void __launch_bounds__(GT)
swarm_grid_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t s_first, uint32_t n_steps, PredSchedule sched,
                  float4 *__restrict__ pos /* [2][gridDim.x * GT] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    cg::grid_group grid = cg::this_grid();
    const int N = S.N, NP = PRED ? S.Npred : 0;
    const int npad = (int)gridDim.x * GT;
    const int n_tiles = (npad + TILE - 1) / TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    RowResult *res_s = reinterpret_cast<RowResult *>(smem_raw);            // [GT]
    float2 *flee_s = reinterpret_cast<float2 *>(res_s + GT);               // [GT]
    double *wred_s = reinterpret_cast<double *>(flee_s + GT);              // [NW][5]
    unsigned *act_s = reinterpret_cast<unsigned *>(wred_s + NW * 5);       // [NW]
    int *n_rows_s = reinterpret_cast<int *>(act_s + NW);                   // [4]
    unsigned short *rows_s = reinterpret_cast<unsigned short *>(n_rows_s + 4);  // [GT]
    float4 *tile_s = reinterpret_cast<float4 *>(rows_s + GT);              // [STAGES][TILE]
    float4 *pred_s = tile_s + STAGES * TILE;                               // [NP]
    float *pphi_s = reinterpret_cast<float *>(pred_s + NP);                // [NP]

    const int t = (int)blockIdx.x * GT + tid;   // agent index
    const bool valid = t < N;
    const size_t g = valid ? (size_t)t : 0;
    AgentState a = {};
    bool alive = false, died = false;
    if (valid) {
        const float4 p = S.pv[g];
        const float4 h = S.hv[g];
        const float2 f = S.force[g];
        const uint2 tm = S.timers[g];
        a.x = p.x; a.y = p.y; a.vx = p.z; a.vy = p.w;
        a.phi = h.x; a.vproj = h.y; a.cu = h.z; a.su = h.w;
        a.fx = f.x; a.fy = f.y;
        a.bin_step = tm.x; a.stb = tm.y;
        alive = S.alive[g] != 0;
    }
    if (PRED) {
        for (int j = tid; j < NP; j += GT) {
            pred_s[j] = S.pred_pv[j];
            pphi_s[j] = S.pred_phi[j];
        }
        __syncthreads();
    }
    uint32_t flags_acc = 0;
    unsigned long long n_amb = 0, n_pairs = 0;
    bool spawned = false;
    const float qnan = __int_as_float(0x7fc00000);

    // CreateFishNet (agents_dynamics.cpp:485-528): centre of mass of the published swarm, computed by every
    // CTA for itself so that all copies of the net are identical
    auto spawn = [&](const float4 *cur, uint32_t step, uint32_t slot) {
        double v[5] = {0, 0, 0, 0, 0};
        const float scaling = SBC_TWO_PI_F * P.invL;
        for (int j = tid; j < N; j += GT) {
            const float4 p = __ldcg(cur + j);
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
        for (int k = 0; k < 5; k++) {
            const double w = warp_sum_d(v[k]);
            if (lane == 0) wred_s[warp * 5 + k] = w;
        }
        __syncthreads();
        for (int k = 0; k < 5; k++) {
            double tot = 0;
            for (int w = 0; w < NW; w++) tot += wred_s[w * 5 + k];
            v[k] = tot;
        }
        __syncthreads();
        float comx, comy;
        if (P.BC == 0) {
            comx = (float)((atan2(-v[1] / v[4], -v[0] / v[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
            comy = (float)((atan2(-v[3] / v[4], -v[2] / v[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
        } else {
            comx = (float)(v[0] / v[4]);
            comy = (float)(v[2] / v[4]);
        }
        const uint4 q = sbc_draw(P, 0, SBC_PREDATOR_AGENT_ID, step, slot);
        const float u0 = (float)((double)q.x * (1.0 / 4294967296.0));
        const float u1 = (float)((double)q.y * (1.0 / 4294967296.0));
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
        for (int j = tid; j < NP; j += GT) {
            float x = sx + px * ((float)j * net_dist), y = sy + py * ((float)j * net_dist);
            float vx = P.pred_speed * cs, vy = P.pred_speed * sn, ph = phi;
            sbc_boundary(x, y, vx, vy, ph, P, P.BC);
            pred_s[j] = make_float4(x, y, vx, vy);
            pphi_s[j] = ph;
        }
        __syncthreads();
    };

    for (uint32_t s = s_first; s != s_first + n_steps; s++) {
        const bool pred_phase = PRED && (s >= sched.s_pred);
        float4 *cur = pos + (size_t)(s & 1u) * npad;
        cur[t] = alive ? make_float4(a.x, a.y, a.vx, a.vy) : make_float4(qnan, qnan, 0.f, 0.f);
        grid.sync();   // the one grid barrier of the step: everybody's positions are in L2
        if (PRED) {
            if (s == sched.s_pred) { spawn(cur, s, 0); spawned = true; }   // swarmdyn.cpp:153-158
            if (pred_phase && sched.needed > 0 && ((int)s - sched.s_trunc) % sched.needed == 0) spawn(cur, s, 1);   // :160-170
        }
        const bool first = alive && (a.bin_step == P.burst_steps);
        const unsigned m = __ballot_sync(FULL, first);
        if (lane == 0) act_s[warp] = m;
        __syncthreads();
        RowAcc acc;
        row_acc_clear(acc);
        if (warp == 0) {  // ordered list of this CTA's burst-starting rows
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
        // The CTA's rows are dealt to its warps (up to RPW rows per warp and pass). All N published positions
        // stream through shared memory in tiles of TILE agents, fetched with cp.async.cg (L2 -> smem, no L1)
        // STAGES - 1 tiles ahead; every warp tests each tile element against its rows.
        for (int r0 = 0; r0 < n_first; r0 += NW * RPW) {
            float4 pi[RPW];
            int row_l[RPW];
            int nr = 0;
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int kr = r0 + warp + q * NW;
                row_l[q] = 0;
                pi[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kr < n_first) {
                    row_l[q] = rows_s[kr];
                    pi[q] = __ldcg(cur + (int)blockIdx.x * GT + row_l[q]);
                    nr = q + 1;
                }
            }
            int cr[RPW], ca[RPW], ct[RPW];
            float rx[RPW], ry[RPW], gx[RPW], gy[RPW], tx[RPW], ty[RPW];
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                cr[q] = ca[q] = ct[q] = 0;
                rx[q] = ry[q] = gx[q] = gy[q] = tx[q] = ty[q] = 0.f;
            }
            auto issue = [&](int k) {
                if (k < n_tiles) {
                    float4 *dst = tile_s + (k % STAGES) * TILE;
                    for (int e = tid; e < TILE; e += GT) {
                        const int j = k * TILE + e;
                        if (j < npad) {
                            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(cur + j) : "memory");
                        } else {
                            dst[e] = make_float4(qnan, qnan, 0.f, 0.f);
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            for (int k = 0; k < STAGES - 1; k++) issue(k);
            for (int k = 0; k < n_tiles; k++) {
                asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
                __syncthreads();   // tile k has landed for everybody; everybody is done with tile k - 1
                issue(k + STAGES - 1);
                const float4 *tl = tile_s + (k % STAGES) * TILE;
                if (nr > 0) {
                    // four tile elements per iteration: their far tests are independent instruction chains; the
                    // zone arithmetic runs only for the few that pass
                    for (int e0 = lane; e0 < TILE; e0 += 128) {
                        float4 pj4[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) pj4[u] = tl[e0 + 32 * u];
#pragma unroll
                        for (int q = 0; q < RPW; q++) {
                            if (q >= nr) break;
                            unsigned near = 0;
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                float dx = pj4[u].x - pi[q].x, dy = pj4[u].y - pi[q].y;
                                if (P.BC == 0) {
                                    dx = sbc_delta_pbc_rough(pi[q].x, pj4[u].x, P);
                                    dy = sbc_delta_pbc_rough(pi[q].y, pj4[u].y, P);
                                }
                                near |= (__fmaf_rn(dy, dy, dx * dx) <= P.far2) ? (1u << u) : 0u;   // false for the dead (NaN)
                            }
                            if (__ballot_sync(FULL, near != 0) == 0) continue;
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                if (!((near >> u) & 1u)) continue;
                                const float4 pj = pj4[u];
                                const int jg = k * TILE + e0 + 32 * u;
                                float dx = pj.x - pi[q].x, dy = pj.y - pi[q].y;
                                if (P.BC == 0) {
                                    dx = sbc_delta_pbc(pi[q].x, pj.x, P);
                                    dy = sbc_delta_pbc(pi[q].y, pj.y, P);
                                }
                                const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                                bool amb;
                                int z = sbc_zone_fast(d2, P, amb);
                                if (amb) {
                                    z = sbc_zone_exact(pi[q].x, pi[q].y, pj.x, pj.y, P);
                                    n_amb++;
                                }
                                if (jg == (int)blockIdx.x * GT + row_l[q]) z = 3;
                                const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
                                if (z == 0) { rx[q] = __fmaf_rn(-dx, rinv, rx[q]); ry[q] = __fmaf_rn(-dy, rinv, ry[q]); cr[q]++; }
                                else if (z == 1) { gx[q] += pj.z; gy[q] += pj.w; ca[q]++; }
                                else if (z == 2) { tx[q] = __fmaf_rn(dx, rinv, tx[q]); ty[q] = __fmaf_rn(dy, rinv, ty[q]); ct[q]++; }
                            }
                        }
                    }
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();   // the stages are free for the next pass
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                if (q >= nr) break;
                const int row = (int)blockIdx.x * GT + row_l[q];
                n_pairs += (lane == 0) ? (unsigned long long)(N - 1) : 0ull;
                const int c0 = __reduce_add_sync(FULL, cr[q]);
                const int c1 = __reduce_add_sync(FULL, ca[q]);
                const int c2 = __reduce_add_sync(FULL, ct[q]);
                float v0 = (c0 > 0) ? rx[q] : gx[q], v1 = (c0 > 0) ? ry[q] : gy[q], v2 = tx[q], v3 = ty[q];
                float f0 = 0.f, f1 = 0.f;
                int cf = 0;
                if (pred_phase) {  // IntCalcPred for this row: a lane takes four net nodes per Philox block
                    RowAcc fa;
                    row_acc_clear(fa);
                    for (int jb = lane; 4 * jb < NP; jb += 32) {
                        const uint4 qd = sbc_draw(P, 0, (uint32_t)row, s, SBC_SLOT_DETECT0 + (uint32_t)jb);
                        const uint32_t r4[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int j = 4 * jb + u;
                            if (j < NP)
                                sbc_detect_accumulate(fa, pi[q].x, pi[q].y, pred_s[j].x, pred_s[j].y,
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
                    r.cr = c0; r.ca = c1; r.ct = c2; r.cf = cf;
                    res_s[row_l[q]] = r;
                    flee_s[row_l[q]] = make_float2(f0, f1);
                }
            }
        }
        __syncthreads();
        if (first) {
            const RowResult r = res_s[tid];
            acc.c_rep = r.cr; acc.c_alg = r.ca; acc.c_att = r.ct; acc.c_flee = r.cf;
            if (r.cr > 0) { acc.rep_x = r.v0; acc.rep_y = r.v1; }
            else { acc.alg_x = r.v0; acc.alg_y = r.v1; }
            acc.att_x = r.v2; acc.att_y = r.v3;
            acc.flee_x = flee_s[tid].x; acc.flee_y = flee_s[tid].y;
        }
        if (alive) {
            double u_social = 0.0;
            float u_dir = 0.f;
            uint32_t k_next = 1;
            const bool rearm = a.stb <= 1;
            if (first || rearm) {
                const uint4 q = sbc_draw(P, 0, (uint32_t)t, s, SBC_SLOT_DECISION);
                u_social = S.inj_social ? S.inj_social[g] : (double)q.x * (1.0 / 4294967296.0);
                u_dir = S.inj_dir ? S.inj_dir[g] : (float)(q.y >> 8) * (1.0f / 16777216.0f);
                if (rearm) k_next = S.inj_k ? S.inj_k[g] : sbc_geometric_k(S.geo, P.max_steps, q.z, P.inv_log2q);
            }
            uint32_t fl = 0;
            sbc_agent_step(a, acc, u_social, u_dir, k_next, P, fl, P.BC);
            flags_acc |= fl;
        }
        if (pred_phase) {  // swarmdyn.cpp:192-203 for a fish net; every CTA keeps an identical copy of the nodes
            __syncthreads();   // the detection above has read the nodes
            for (int j = tid; j < NP; j += GT) {  // MoveFishNet :616-624
                float4 p = pred_s[j];
                float ph = pphi_s[j];
                p.x += p.z * P.dt;
                p.y += p.w * P.dt;
                sbc_boundary(p.x, p.y, p.z, p.w, ph, P, P.BC);
                pred_s[j] = p;
                pphi_s[j] = ph;
            }
            __syncthreads();
            if (P.pred_kill != 0 && alive && sbc_net_kills(a.x, a.y, pred_s[0], pred_s[1], pphi_s[0], NP, P)) {  // FishNetKill
                alive = false;
                died = true;
            }
        }
    }
    if (valid && (alive || died)) {
        S.pv[g] = make_float4(a.x, a.y, a.vx, a.vy);
        S.hv[g] = make_float4(a.phi, a.vproj, a.cu, a.su);
        S.force[g] = make_float2(a.fx, a.fy);
        S.timers[g] = make_uint2(a.bin_step, a.stb);
        if (S.flags) S.flags[g] |= (uint8_t)flags_acc;
        if (died) {  // split_dead (agents_operation.cpp:209-218)
            S.alive[g] = 0;
            S.pv_dead[g] = make_float4(a.x, a.y, a.vx, a.vy);
            S.pv[g] = make_float4(qnan, qnan, a.vx, a.vy);
            atomicAdd(S.n_dead, 1);
        }
    }
    if (PRED && blockIdx.x == 0) {
        __syncthreads();
        for (int j = tid; j < NP; j += GT) {
            S.pred_pv[j] = pred_s[j];
            S.pred_phi[j] = pphi_s[j];
        }
        if (tid == 0 && spawned) S.pred_spawned[0] = 1;
    }
    if (n_amb) atomicAdd(S.stats + 4, n_amb);
    if (n_pairs) atomicAdd(S.stats + 3, n_pairs);
}

size_t grid_smem_bytes(int npred) {
    return GT * sizeof(RowResult) + GT * sizeof(float2) + NW * 5 * sizeof(double) + NW * sizeof(unsigned) + 4 * sizeof(int) +
           GT * sizeof(unsigned short) + (size_t)STAGES * TILE * sizeof(float4) + (size_t)npred * (sizeof(float4) + sizeof(float)) + 64;
}

int g_max_ctas[2] = {-1, -1};   // co-resident CTAs of the two instantiations on this device

}  // namespace

// a single swarm, no single predator (fish nets are fine), metric neighbour rule, all CTAs co-resident
bool sbc_grid_supported(const DevParams &P, int N, int R, int Npred) {
    if (R != 1 || P.voronoi || Npred == 1 || Npred > 2048 || N <= GT) return false;
    const int pred = Npred > 0 ? 1 : 0;
    if (g_max_ctas[pred] < 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t smem = grid_smem_bytes(pred ? 2048 : 0);
        if (pred) {
            cudaFuncSetAttribute(swarm_grid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, swarm_grid_kernel<true>, GT, smem);
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, swarm_grid_kernel<false>, GT, smem);
        }
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        g_max_ctas[pred] = coop ? sms * per_sm : 0;
    }
    return (N + GT - 1) / GT <= g_max_ctas[pred];
}

cudaError_t sbc_launch_swarm_grid(const DevParams &P, const DevState &S, long long s_first, long long n_steps, float4 *pos,
                                  cudaStream_t st, long long s_pred, int s_trunc, int needed) {
    const int C = (S.N + GT - 1) / GT;
    PredSchedule sched;
    sched.s_pred = (s_pred < 0 || s_pred > 0xffffffffll) ? 0xffffffffu : (uint32_t)s_pred;
    sched.s_trunc = s_trunc;
    sched.needed = needed;
    uint32_t s0 = (uint32_t)s_first, ns = (uint32_t)n_steps;
    DevParams Pc = P;
    DevState Sc = S;
    void *args[] = {&Pc, &Sc, &s0, &ns, &sched, &pos};
    const size_t smem = grid_smem_bytes(S.Npred);
    if (S.Npred > 0)
        return cudaLaunchCooperativeKernel((const void *)swarm_grid_kernel<true>, dim3((unsigned)C), dim3(GT), args, smem, st);
    return cudaLaunchCooperativeKernel((const void *)swarm_grid_kernel<false>, dim3((unsigned)C), dim3(GT), args, smem, st);
}
