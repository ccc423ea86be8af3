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
    const uint4 q = sbc_draw(P, rep, SBC_PREDATOR_AGENT_ID, step, slot);
    const float u0 = (float)((double)q.x * (1.0 / 4294967296.0));
    const float u1 = (float)((double)q.y * (1.0 / 4294967296.0));
    float4 *pp = S.pred_pv + (size_t)rep * S.Npred;
    float *pphi = S.pred_phi + (size_t)rep * S.Npred;
    if (S.Npred == 1) {
        if (threadIdx.x == 0) {
            float phi = u0 * SBC_TWO_PI_F, s, c;
            sincosf(phi, &s, &c);
            float x = s_com[0] - 0.5f * P.L * c, y = s_com[1] - 0.5f * P.L * s;
            float vx = P.pred_speed * c, vy = P.pred_speed * s;
            sbc_boundary(x, y, vx, vy, phi, P, P.BC);
            pp[0] = make_float4(x, y, vx, vy);
            pphi[0] = phi;
            S.pred_spawned[rep] = 1;
        }
    } else {
        float phi = SBC_TWO_PI_F * u0, s, c;
        sincosf(phi, &s, &c);
        float sx = s_com[0] - 0.25f * P.L * c, sy = s_com[1] - 0.25f * P.L * s;
        const float var_noise = SBC_PI_F / 4.f;
        phi += var_noise * u1 - 0.5f * var_noise;
        sincosf(phi, &s, &c);
        const float px = s, py = -c;  // vec_perp(v_net_move), mathtools.h:183-189
        const float net_len = 0.25f * P.L;
        sx -= 0.5f * net_len * px;
        sy -= 0.5f * net_len * py;
        const float net_dist = net_len / (float)(S.Npred - 1);
        for (int j = threadIdx.x; j < S.Npred; j += PT) {
            float x = sx + px * ((float)j * net_dist), y = sy + py * ((float)j * net_dist);
            float vx = P.pred_speed * c, vy = P.pred_speed * s, ph = phi;
            sbc_boundary(x, y, vx, vy, ph, P, P.BC);
            pp[j] = make_float4(x, y, vx, vy);
            pphi[j] = ph;
        }
        if (threadIdx.x == 0) S.pred_spawned[rep] = 1;
    }
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
