// predator.cu - predator / fish-net spawn, move and kill for swarms on the multi-kernel path.
//   spawn : GetCenterOfMass (agents_operation.cpp:23-127) + CreatePredator (agents_dynamics.cpp:
//           534-555) or CreateFishNet (:485-528)
//   move  : MovePredator (:567-613, random walk or chase of the closest prey) / MoveFishNet (:616-624)
//   kill  : PredKill (:826-915, pred_kill 1..4) / FishNetKill (:780-820)
// One thread block per replicate; the per-agent kill predicates are evaluated in FP64 from the
// FP32 state so that dead sets follow the reference's comparisons exactly for a given state.
#include "sbc_kernels.cuh"

namespace {

constexpr int PT = 256;

__device__ __forceinline__ double block_sum(double v, double *sh) {
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < PT / 32; w++) t += sh[w];
    return t;
}

// CreatePredator (agents_dynamics.cpp:534-555) / CreateFishNet (:485-528) around the swarm's centre of mass, by the whole block
__device__ void pred_create(const DevParams &P, const DevState &S, int rep, uint32_t step, uint32_t slot, float com0, float com1) {
    const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, step, slot);
    const float u0 = (float)((double)q.x * (1.0 / 4294967296.0));
    const float u1 = (float)((double)q.y * (1.0 / 4294967296.0));
    float4 *pp = S.pred_pv + (size_t)rep * S.Npred;
    float *pphi = S.pred_phi + (size_t)rep * S.Npred;
    if (S.Npred == 1) {
        if (threadIdx.x == 0) {
            float phi = u0 * SBC_TWO_PI_F, s, c;
            sincosf(phi, &s, &c);
            float x = com0 - 0.5f * P.L * c, y = com1 - 0.5f * P.L * s;
            float vx = P.pred_speed * c, vy = P.pred_speed * s;
            sbc_boundary(x, y, vx, vy, phi, P, P.BC);
            pp[0] = make_float4(x, y, vx, vy);
            pphi[0] = phi;
            S.pred_spawned[rep] = 1;
        }
    } else {
        float phi = SBC_TWO_PI_F * u0, s, c;
        sincosf(phi, &s, &c);
        float sx = com0 - 0.25f * P.L * c, sy = com1 - 0.25f * P.L * s;
        const float var_noise = SBC_PI_F / 4.f;
        phi += var_noise * u1 - 0.5f * var_noise;
        sincosf(phi, &s, &c);
        const float px = s, py = -c;  // vec_perp(v_net_move), mathtools.h:183-189
        const float net_len = 0.25f * P.L;
        sx -= 0.5f * net_len * px;
        sy -= 0.5f * net_len * py;
        const float net_dist = net_len / (float)(S.Npred - 1);
        for (int j = threadIdx.x; j < S.Npred; j += blockDim.x) {
            float x = sx + px * ((float)j * net_dist), y = sy + py * ((float)j * net_dist);
            float vx = P.pred_speed * c, vy = P.pred_speed * s, ph = phi;
            sbc_boundary(x, y, vx, vy, ph, P, P.BC);
            pp[j] = make_float4(x, y, vx, vy);
            pphi[j] = ph;
        }
        if (threadIdx.x == 0) S.pred_spawned[rep] = 1;
    }
}

__global__ void __launch_bounds__(PT)
pred_spawn_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t step, uint32_t slot) {
    __shared__ double sh[PT / 32];
    __shared__ float s_com[2];
    const int rep = blockIdx.x;
    const int N = S.N;
    const float4 *pv = S.pv + (size_t)rep * N;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, cnt = 0;
    const float scaling = SBC_TWO_PI_F * P.invL;
    for (int i = threadIdx.x; i < N; i += PT) {
        const float4 p = pv[i];
        if (p.x != p.x) continue;
        cnt += 1;
        if (P.BC == 0) {
            float s, c;
            sincosf(scaling * p.x, &s, &c);
            a0 += c; a1 += s;
            sincosf(scaling * p.y, &s, &c);
            a2 += c; a3 += s;
        } else {
            a0 += p.x;
            a2 += p.y;
        }
    }
    a0 = block_sum(a0, sh); a1 = block_sum(a1, sh);
    a2 = block_sum(a2, sh); a3 = block_sum(a3, sh);
    cnt = block_sum(cnt, sh);
    if (threadIdx.x == 0) {
        if (P.BC == 0) {
            s_com[0] = (float)((atan2(-a1 / cnt, -a0 / cnt) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
            s_com[1] = (float)((atan2(-a3 / cnt, -a2 / cnt) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
        } else {
            s_com[0] = (float)(a0 / cnt);
            s_com[1] = (float)(a2 / cnt);
        }
    }
    __syncthreads();
    pred_create(P, S, rep, step, slot, s_com[0], s_com[1]);
}

// MoveFishNet (agents_dynamics.cpp:616-624): one block per replicate; block 0 first applies the kills the step's
// kernels recorded (FishNetKill was evaluated per agent against the moved net, sbc_update_agent): S.pv holds the
// positions after this step's update here
__global__ void __launch_bounds__(PT)
net_move_kernel(const __grid_constant__ DevParams P, DevState S) {
    const int rep = blockIdx.x;
    if (blockIdx.x == 0 && S.kill_count) {
        const int nk = *S.kill_count;
        for (int k = threadIdx.x; k < nk; k += PT) {
            const size_t g = (size_t)S.kill_list[k];
            sbc_kill_agent(S, g, S.pv[g]);
            atomicAdd(S.n_dead + g / S.N, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) *S.kill_count = 0;
    }
    float4 *pp = S.pred_pv + (size_t)rep * S.Npred;
    float *pphi = S.pred_phi + (size_t)rep * S.Npred;
    for (int j = threadIdx.x; j < S.Npred; j += PT) {
        float4 p = pp[j];
        float ph = pphi[j];
        p.x += p.z * P.dt;
        p.y += p.w * P.dt;
        sbc_boundary(p.x, p.y, p.z, p.w, ph, P, P.BC);
        pp[j] = p;
        pphi[j] = ph;
    }
}

__global__ void __launch_bounds__(PT)
pred_move_kill_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t step_arg,
                      const uint32_t *__restrict__ step_dev) {
    const uint32_t step = step_dev ? *step_dev + step_arg : step_arg;   // graph-replayed: base + offset inside the launch
    __shared__ double sh[PT / 32];
    __shared__ unsigned long long s_key[PT / 32];
    __shared__ float s_p[4];
    __shared__ int s_killed;
    const int rep = blockIdx.x;
    const int N = S.N;
    float4 *pv = S.pv + (size_t)rep * N;
    float4 *pp = S.pred_pv + (size_t)rep * S.Npred;
    float *pphi = S.pred_phi + (size_t)rep * S.Npred;
    if (threadIdx.x == 0) s_killed = 0;
    if (S.Npred > 1) {
        // fish nets are handled by net_move_kernel + net_kill_kernel (grid over the agents)
    } else {
        // MovePredator
        const float4 p0 = pp[0];
        float lphi;
        if (P.pred_move == 0) {
            const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, step, 2u);
            const float u1 = (float)(((double)q.x + 1.0) * (1.0 / 4294967296.0));
            const float u2 = (float)((double)q.y * (1.0 / 4294967296.0));
            const float gss = sqrtf(-2.f * logf(u1)) * cosf(SBC_TWO_PI_F * u2);
            lphi = fmodf(pphi[0] + P.noisep * gss * sqrtf(P.pred_speed), SBC_TWO_PI_F);
        } else {
            unsigned long long best = ~0ull;
            for (int i = threadIdx.x; i < N; i += PT) {
                const float4 p = pv[i];
                if (p.x != p.x) continue;
                float dx = p.x - p0.x, dy = p.y - p0.y;
                if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, p.x, P); dy = sbc_delta_pbc(p0.y, p.y, P); }
                const float d2 = dx * dx + dy * dy;
                const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i;
                best = min(best, key);
            }
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
            __syncthreads();
            best = ~0ull;
            for (int w = 0; w < PT / 32; w++) best = min(best, s_key[w]);
            if (best == ~0ull) {
                lphi = pphi[0];
            } else {
                const float4 p = pv[(unsigned)(best & 0xffffffffu)];
                float dx = p.x - p0.x, dy = p.y - p0.y;
                if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, p.x, P); dy = sbc_delta_pbc(p0.y, p.y, P); }
                lphi = atan2f(dy, dx);
            }
        }
        if (threadIdx.x == 0) {
            float s, c;
            sincosf(lphi, &s, &c);
            float vx = P.pred_speed * c, vy = P.pred_speed * s;
            float x = p0.x + vx * P.dt, y = p0.y + vy * P.dt;
            float ph = lphi;
            sbc_boundary(x, y, vx, vy, ph, P, P.BC);
            pp[0] = make_float4(x, y, vx, vy);
            pphi[0] = ph;
            s_p[0] = x; s_p[1] = y;
        }
        __syncthreads();
        if (P.pred_kill != 0) {  // PredKill
            const float px = s_p[0], py = s_p[1];
            double n_sensed = 0, total = 0;
            // pass 1: N_sensed and the normalisation of the selection weights (:861-877)
            if (P.pred_kill > 1) {
                for (int i = threadIdx.x; i < N; i += PT) {
                    const float4 p = pv[i];
                    if (p.x != p.x) continue;
                    const KillClass kc = sbc_kill_classify(p.x, p.y, px, py, P);
                    if (kc.sensed) {
                        n_sensed += 1;
                        if (kc.cand && P.pred_kill == 4) total += sbc_kill_weight(p.x, p.y, px, py, P);
                    }
                }
                n_sensed = block_sum(n_sensed, sh);
                total = block_sum(total, sh);
            }
            int killed = 0;
            for (int i = threadIdx.x; i < N; i += PT) {
                const float4 p = pv[i];
                if (p.x != p.x) continue;
                const KillClass kc = sbc_kill_classify(p.x, p.y, px, py, P);
                bool die = false;
                if (P.pred_kill == 1) {
                    die = kc.within;
                } else if (kc.cand) {
                    const double pk = sbc_kill_probability(p.x, p.y, px, py, (int)n_sensed, total, S.kill_eff, P);
                    const uint32_t id = S.gid ? S.gid[(size_t)rep * N + i] : (uint32_t)i;
                    const uint4 q = sbc_draw(P, rep, id, step, SBC_SLOT_DECISION);
                    die = pk > (double)q.w * (1.0 / 4294967296.0);
                }
                if (die) {
                    sbc_kill_agent(S, (size_t)rep * N + i, p);
                    killed++;
                }
            }
            if (killed) atomicAdd(&s_killed, killed);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_killed) S.n_dead[rep] += s_killed;
}


// ---- predators of a DECOMPOSED swarm (one x-slab per GPU, domain.cu) ---------------------------------------------------
// The predator state is replicated: every rank holds the same predator and moves it with the same arithmetic. What
// needs all agents - the closest prey of the chase (MovePredator, agents_dynamics.cpp:580-597), the number of sensed
// prey and the weight sum of PredKill (:861-877), the centre of mass at spawn (agents_operation.cpp:23-127) - travels
// as ONE 64-byte record per rank and round through peer memory: every rank stores its record into the predator area
// of every rank's inbox, raises a flag there, and waits for all flags of the round in its own area (an all-gather of
// world x 64 bytes; no message library). Records are combined in rank order, so every rank computes the same result and
// the result does not depend on timing. Round numbers are counted on the device (ctl[24 + round]), two buffers per
// round by parity: a rank can only be one round ahead of the slowest, because finishing a round needs everybody's record.
enum { PB_ARGMIN = 0, PB_KILL = 1, PB_SPAWN = 2, PB_ROUNDS = 3, PB_MAX_WORLD = 16, PB_REC = 64 };
struct PbRec { double v[8]; };
__device__ __forceinline__ PbRec *pb_rec(unsigned char *area, int round, int par, int r) {
    return reinterpret_cast<PbRec *>(area + (size_t)((round * 2 + par) * PB_MAX_WORLD + r) * PB_REC);
}
__device__ __forceinline__ int *pb_flag(unsigned char *area, int round, int par, int r) {
    return reinterpret_cast<int *>(area + (size_t)PB_ROUNDS * 2 * PB_MAX_WORLD * PB_REC) + (round * 2 + par) * PB_MAX_WORLD + r;
}
// called by all 32 lanes of warp 0; out[r] (shared memory) = the record of rank r. false: a peer did not answer in time
__device__ bool pb_allgather(const PredBox &B, int *ctl, int round, const PbRec &mine, PbRec *out) {
    const int lane = threadIdx.x & 31;
    int tag = 0;
    if (lane == 0) tag = ++ctl[24 + round];
    tag = __shfl_sync(0xffffffffu, tag, 0);
    const int par = tag & 1;
    bool ok = true;
    if (lane < B.world) {
        unsigned char *peer = B.peer[0];
#pragma unroll
        for (int r = 1; r < PB_MAX_WORLD; r++) if (r == lane) peer = B.peer[r];   // selects: no local copy of the parameter
        unsigned char *own = B.peer[0];
#pragma unroll
        for (int r = 1; r < PB_MAX_WORLD; r++) if (r == B.rank) own = B.peer[r];
        volatile double *dst = pb_rec(peer, round, par, B.rank)->v;
#pragma unroll
        for (int k = 0; k < 8; k++) dst[k] = mine.v[k];
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(pb_flag(peer, round, par, B.rank)), "r"(tag) : "memory");
        const int *f = pb_flag(own, round, par, lane);
        const unsigned long long t0 = sbc_now_ns();
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v >= tag) break;
            if (ctl[21] || sbc_now_ns() - t0 > B.timeout_ns) { ok = false; break; }
            __nanosleep(40);
        }
        const volatile double *src = pb_rec(own, round, par, lane)->v;
#pragma unroll
        for (int k = 0; k < 8; k++) out[lane].v[k] = src[k];
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok && lane == 0 && ctl[21] == 0) { ctl[6] += 1000; ctl[21] = 1; }   // same record as a failed halo exchange
    __syncwarp();
    return ok;
}

// a killed agent that is mirrored on a neighbour tells it so: the refresh record of this step already left with the
// agent's position (sbc_update_agent delivers it), the notice overwrites it before the flags go up
__device__ __forceinline__ void dom_death_notice(const StepArgs &A, size_t g, uint32_t step) {
    if (!A.halo_idx) return;
    const int2 hi = A.halo_idx[g];
    if ((hi.x & hi.y) == -1) return;
    const int par = ((int)step + A.xshift[0]) & 1;
    const float qnan = __int_as_float(0x7fc00000);
    const float4 npv = make_float4(qnan, qnan, 0.f, 0.f);
    void *dl = par ? A.xdst[0][1] : A.xdst[0][0], *dr = par ? A.xdst[1][1] : A.xdst[1][0];
    if (hi.x >= 0) reinterpret_cast<float4 *>(dl)[(size_t)(1 + hi.x) * 4] = npv;
    if (hi.y >= 0) reinterpret_cast<float4 *>(dr)[(size_t)(1 + hi.y) * 4] = npv;
}

// spawn: partial sums of GetCenterOfMass over the owned slots, all-gather, CreatePredator / CreateFishNet everywhere
__global__ void __launch_bounds__(PT)
dom_pred_spawn_kernel(const __grid_constant__ DevParams P, DevState S, const __grid_constant__ PredBox B, int own_cap, int *ctl,
                      uint32_t step, uint32_t slot) {
    __shared__ double sh[PT / 32];
    __shared__ PbRec recs[PB_MAX_WORLD];
    __shared__ float s_com[2];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, cnt = 0;
    const float scaling = SBC_TWO_PI_F * P.invL;
    for (int i = threadIdx.x; i < own_cap; i += PT) {
        const float4 p = S.pv[i];
        if (p.x != p.x) continue;
        cnt += 1;
        float s, c;
        sincosf(scaling * p.x, &s, &c);
        a0 += c; a1 += s;
        sincosf(scaling * p.y, &s, &c);
        a2 += c; a3 += s;
    }
    a0 = block_sum(a0, sh); a1 = block_sum(a1, sh); a2 = block_sum(a2, sh); a3 = block_sum(a3, sh); cnt = block_sum(cnt, sh);
    if (threadIdx.x < 32) {
        PbRec mine = {{a0, a1, a2, a3, cnt, 0, 0, 0}};
        pb_allgather(B, ctl, PB_SPAWN, mine, recs);
        if (threadIdx.x == 0) {
            double t[5] = {0, 0, 0, 0, 0};
            for (int r = 0; r < B.world; r++) for (int k = 0; k < 5; k++) t[k] += recs[r].v[k];
            s_com[0] = (float)((atan2(-t[1] / t[4], -t[0] / t[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
            s_com[1] = (float)((atan2(-t[3] / t[4], -t[2] / t[4]) + SBC_PI_D) / (2 * SBC_PI_D / P.dL));
        }
    }
    __syncthreads();
    pred_create(P, S, 0, step, slot, s_com[0], s_com[1]);
}

// chase, pass 1: this rank's closest prey of the predator, key = (FP32 squared distance, slot)
__global__ void __launch_bounds__(PT)
dom_pred_argmin_kernel(const __grid_constant__ DevParams P, DevState S, int own_cap, unsigned long long *key_out) {
    const float4 p0 = S.pred_pv[0];
    unsigned long long best = ~0ull;
    for (int i = blockIdx.x * PT + threadIdx.x; i < own_cap; i += gridDim.x * PT) {
        const float4 p = S.pv[i];
        if (p.x != p.x) continue;
        float dx = p.x - p0.x, dy = p.y - p0.y;
        if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, p.x, P); dy = sbc_delta_pbc(p0.y, p.y, P); }
        const float d2 = dx * dx + dy * dy;
        best = min(best, ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i);
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(key_out, best);
}

// MovePredator on every rank (one warp): all-gather of the ranks' closest prey (chase) or the random-walk draw
__global__ void __launch_bounds__(32)
dom_pred_move_kernel(const __grid_constant__ DevParams P, DevState S, const __grid_constant__ PredBox B, int *ctl,
                     unsigned long long *key, StepArgs A) {
    __shared__ PbRec recs[PB_MAX_WORLD];
    const uint32_t step = sbc_step_of(A);
    const float4 p0 = S.pred_pv[0];
    float lphi = S.pred_phi[0];
    if (P.pred_move == 0) {
        const uint4 q = sbc_draw(P, 0, SBC_PREDATOR_AGENT_ID, step, 2u);
        const float u1 = (float)(((double)q.x + 1.0) * (1.0 / 4294967296.0));
        const float u2 = (float)((double)q.y * (1.0 / 4294967296.0));
        const float gss = sqrtf(-2.f * logf(u1)) * cosf(SBC_TWO_PI_F * u2);
        lphi = fmodf(S.pred_phi[0] + P.noisep * gss * sqrtf(P.pred_speed), SBC_TWO_PI_F);
    } else {
        const unsigned long long k = *key;
        PbRec mine = {{INFINITY, 0, 0, 0, 0, 0, 0, 0}};
        if (k != ~0ull) {
            const unsigned slot = (unsigned)(k & 0xffffffffu);
            const float4 p = S.pv[slot];
            mine.v[0] = (double)__uint_as_float((unsigned)(k >> 32));
            mine.v[1] = (double)(S.gid ? S.gid[slot] : slot);
            mine.v[2] = p.x; mine.v[3] = p.y;
        }
        pb_allgather(B, ctl, PB_ARGMIN, mine, recs);
        int win = -1;   // smallest (distance, global id): the key order of the undecomposed kernel
        for (int r = 0; r < B.world; r++)
            if (recs[r].v[0] < INFINITY &&
                (win < 0 || recs[r].v[0] < recs[win].v[0] || (recs[r].v[0] == recs[win].v[0] && recs[r].v[1] < recs[win].v[1])))
                win = r;
        if (win >= 0) {
            const float wx = (float)recs[win].v[2], wy = (float)recs[win].v[3];
            float dx = wx - p0.x, dy = wy - p0.y;
            if (P.BC == 0) { dx = sbc_delta_pbc(p0.x, wx, P); dy = sbc_delta_pbc(p0.y, wy, P); }
            lphi = atan2f(dy, dx);
        }
    }
    if (threadIdx.x == 0) {
        float s, c;
        sincosf(lphi, &s, &c);
        float vx = P.pred_speed * c, vy = P.pred_speed * s;
        float x = p0.x + vx * P.dt, y = p0.y + vy * P.dt;
        float ph = lphi;
        sbc_boundary(x, y, vx, vy, ph, P, P.BC);
        S.pred_pv[0] = make_float4(x, y, vx, vy);
        S.pred_phi[0] = ph;
        *key = ~0ull;
        ctl[27] = 0;   // sensed prey of this rank
        ctl[28] = 0;   // kill candidates of this rank
    }
}

__device__ __forceinline__ void dom_kill(const DevState &S, const StepArgs &A, size_t g, float4 p, uint32_t step) {
    sbc_kill_agent(S, g, p);
    dom_death_notice(A, g, step);
    atomicAdd(S.n_dead, 1);
}

// PredKill, pass 1 over the owned slots against the MOVED predator: rules 1 and 2 need nothing global and kill here;
// rules 3 and 4 count the sensed prey and list the candidates (d < kill_range) for dom_pred_kill_kernel
__global__ void __launch_bounds__(PT)
dom_pred_sense_kernel(const __grid_constant__ DevParams P, DevState S, int own_cap, int *ctl, int *cand, StepArgs A) {
    const uint32_t step = sbc_step_of(A);
    const float px = S.pred_pv[0].x, py = S.pred_pv[0].y;
    for (int i = blockIdx.x * PT + threadIdx.x; i < own_cap; i += gridDim.x * PT) {
        const float4 p = S.pv[i];
        if (p.x != p.x) continue;
        const KillClass kc = sbc_kill_classify(p.x, p.y, px, py, P);
        if (P.pred_kill == 1) {
            if (kc.within) dom_kill(S, A, (size_t)i, p, step);
        } else if (P.pred_kill == 2) {
            if (kc.cand) {
                const double pk = sbc_kill_probability(p.x, p.y, px, py, 0, 0.0, S.kill_eff, P);
                const uint4 q = sbc_draw(P, 0, S.gid ? S.gid[i] : (uint32_t)i, step, SBC_SLOT_DECISION);
                if (pk > (double)q.w * (1.0 / 4294967296.0)) dom_kill(S, A, (size_t)i, p, step);
            }
        } else if (kc.sensed) {
            atomicAdd(ctl + 27, 1);
            if (kc.cand) cand[atomicAdd(ctl + 28, 1)] = i;
        }
    }
}

// PredKill rules 3 and 4 (one block): all-gather of (sensed prey, weight sum), then the candidates' draws
__global__ void __launch_bounds__(PT)
dom_pred_kill_kernel(const __grid_constant__ DevParams P, DevState S, const __grid_constant__ PredBox B, int *ctl, int *cand,
                     StepArgs A) {
    __shared__ PbRec recs[PB_MAX_WORLD];
    __shared__ double s_tot[2];
    const uint32_t step = sbc_step_of(A);
    const float px = S.pred_pv[0].x, py = S.pred_pv[0].y;
    const int nc = ctl[28];
    if (threadIdx.x < 32) {
        double total = 0;
        if (threadIdx.x == 0 && P.pred_kill == 4) {
            // the weights are summed in the order of the global ids, whatever order the candidates were listed in
            for (int a = 1; a < nc; a++) {
                const int v = cand[a];
                const uint32_t gv = S.gid ? S.gid[v] : (uint32_t)v;
                int b = a - 1;
                while (b >= 0 && (S.gid ? S.gid[cand[b]] : (uint32_t)cand[b]) > gv) { cand[b + 1] = cand[b]; b--; }
                cand[b + 1] = v;
            }
            for (int a = 0; a < nc; a++) { const float4 p = S.pv[cand[a]]; total += sbc_kill_weight(p.x, p.y, px, py, P); }
        }
        PbRec mine = {{(double)ctl[27], total, 0, 0, 0, 0, 0, 0}};
        pb_allgather(B, ctl, PB_KILL, mine, recs);
        if (threadIdx.x == 0) {
            double ns = 0, tt = 0;
            for (int r = 0; r < B.world; r++) { ns += recs[r].v[0]; tt += recs[r].v[1]; }
            s_tot[0] = ns; s_tot[1] = tt;
        }
    }
    __syncthreads();
    const int n_sensed = (int)s_tot[0];
    const double total = s_tot[1];
    for (int a = threadIdx.x; a < nc; a += PT) {
        const int i = cand[a];
        const float4 p = S.pv[i];
        const double pk = sbc_kill_probability(p.x, p.y, px, py, n_sensed, total, S.kill_eff, P);
        const uint4 q = sbc_draw(P, 0, S.gid ? S.gid[i] : (uint32_t)i, step, SBC_SLOT_DECISION);
        if (pk > (double)q.w * (1.0 / 4294967296.0)) dom_kill(S, A, (size_t)i, p, step);
    }
}

}  // namespace

void sbc_launch_pred_spawn(const DevParams &P, const DevState &S, long long step, int slot,
                           cudaStream_t st) {
    pred_spawn_kernel<<<S.R, PT, 0, st>>>(P, S, (uint32_t)step, (uint32_t)slot);
}
void sbc_launch_pred_move_kill(const DevParams &P, const DevState &S, long long step, const uint32_t *step_dev,
                               cudaStream_t st) {
    if (S.Npred > 1) {  // the net kill was evaluated by update_kernel against the moved first two nodes
        net_move_kernel<<<S.R, PT, 0, st>>>(P, S);
        return;
    }
    pred_move_kill_kernel<<<S.R, PT, 0, st>>>(P, S, (uint32_t)step, step_dev);
}

// ---- decomposed swarm (PredBox: the predator areas of every rank's inbox) ----
// Load the kernels of the predator phase NOW. With lazy module loading the first launch of a kernel loads it, which can
// wait for every kernel running on the device - including another handle's all-gather that is waiting for this very
// handle's record (several slabs on one GPU). Called when the predator areas are attached, before anything waits.
void sbc_pred_preload(void) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, dom_pred_spawn_kernel);
    cudaFuncGetAttributes(&fa, dom_pred_argmin_kernel);
    cudaFuncGetAttributes(&fa, dom_pred_move_kernel);
    cudaFuncGetAttributes(&fa, dom_pred_sense_kernel);
    cudaFuncGetAttributes(&fa, dom_pred_kill_kernel);
    cudaFuncGetAttributes(&fa, net_move_kernel);
    sbc_cells_preload();
}
size_t sbc_pred_area_bytes(void) { return 8192; }   // PB_ROUNDS * 2 * PB_MAX_WORLD * (PB_REC + 4) = 6528, rounded up
void sbc_launch_domain_pred_spawn(const DevParams &P, const DevState &S, const PredBox &B, int own_cap, int *ctl, long long step,
                                  int slot, cudaStream_t st) {
    dom_pred_spawn_kernel<<<1, PT, 0, st>>>(P, S, B, own_cap, ctl, (uint32_t)step, (uint32_t)slot);
}
// the predator phase behind the update of one step; S = the state AFTER the update (record_step's Sn)
void sbc_launch_domain_pred_move_kill(const DevParams &P, const DevState &S, const PredBox &B, int own_cap, int *ctl,
                                      unsigned long long *key, int *cand, const StepArgs &A, cudaStream_t st) {
    if (S.Npred > 1) {   // fish net: moves rigidly, FishNetKill was evaluated by the update (death notices included)
        net_move_kernel<<<1, PT, 0, st>>>(P, S);
        return;
    }
    const int grid = min(148 * 4, (own_cap + PT - 1) / PT);
    if (P.pred_move != 0) dom_pred_argmin_kernel<<<grid, PT, 0, st>>>(P, S, own_cap, key);
    dom_pred_move_kernel<<<1, 32, 0, st>>>(P, S, B, ctl, key, A);
    if (P.pred_kill == 0) return;
    dom_pred_sense_kernel<<<grid, PT, 0, st>>>(P, S, own_cap, ctl, cand, A);
    if (P.pred_kill > 2) dom_pred_kill_kernel<<<1, PT, 0, st>>>(P, S, B, ctl, cand, A);
}
