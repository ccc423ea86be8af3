// update.cu - per-agent burst-coast update for swarms that do not fit one thread block:
// ParticleBurstCoast with the burst decision, wall avoidance, Euler step, boundary and burst
// re-arm (/root/reference/src/agents_dynamics.cpp:150-273). For rows that start a burst while a
// predator is present, the prey<-predator detection of IntCalcPred (agents_interact.cpp:252-292; only
// rows that read the result evaluate it) has been evaluated by the pair kernel that swept
// the row (sbc_detect_row_warp, sbc_kernels.cuh).
//
// One thread per agent, coalesced SoA traffic: 48 B read + 48 B written per agent-step
// (pv float4, hv float4, force float2, timers uint2) -> HBM-bound.
#include "sbc_kernels.cuh"

namespace {

__global__ void __launch_bounds__(256)
update_kernel(const __grid_constant__ DevParams P, DevState S, uint32_t step_arg, const uint32_t *__restrict__ step_dev,
              int pred_active, int *zero_count, int net_kill) {
    // the pair pass and the detection of this step are done: the next compaction may start from zero
    if (zero_count && blockIdx.x == 0 && threadIdx.x == 0) *zero_count = 0;
    // inside a captured CUDA graph the step index lives in device memory (advanced by step_advance_kernel)
    const uint32_t step = step_dev ? *step_dev : step_arg;
    const size_t total = (size_t)S.N * S.R;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    if (!S.alive[g]) return;
    const int N = S.N;
    const uint32_t rep = (uint32_t)(g / N);
    // Philox counters use the GLOBAL agent id, so a decomposed swarm draws what the whole swarm would
    const uint32_t i = S.gid ? S.gid[g] : (uint32_t)(g - (size_t)rep * N);
    AgentState a;
    {
        const float4 p = S.pv[g];
        const float4 h = S.hv[g];
        const float2 f = S.force[g];
        const uint2 t = S.timers[g];
        a.x = p.x; a.y = p.y; a.vx = p.z; a.vy = p.w;
        a.phi = h.x; a.vproj = h.y; a.cu = h.z; a.su = h.w;
        a.fx = f.x; a.fy = f.y;
        a.bin_step = t.x; a.stb = t.y;
    }
    const bool first = (a.bin_step == P.burst_steps);
    const bool rearm = (a.stb <= 1);
    RowAcc acc;
    row_acc_clear(acc);
    double u_social = 0.0;
    float u_dir = 0.f;
    uint32_t k_next = 1;
    if (first || rearm) {
        const uint4 q = sbc_draw(P, rep, i, step, SBC_SLOT_DECISION);
        u_social = S.inj_social ? S.inj_social[g] : (double)q.x * (1.0 / 4294967296.0);
        u_dir = S.inj_dir ? S.inj_dir[g] : (float)(q.y >> 8) * (1.0f / 16777216.0f);
        if (rearm) k_next = S.inj_k ? S.inj_k[g] : sbc_geometric_k(S.geo, P.max_steps, q.z, P.inv_log2q);
    }
    if (first) {
        const float4 *ac = reinterpret_cast<const float4 *>(S.acc + g * 8);
        const float4 a0 = ac[0], a1 = ac[1];
        const int4 c = *reinterpret_cast<const int4 *>(S.cnt + g * 4);
        acc.rep_x = a0.x; acc.rep_y = a0.y; acc.alg_x = a0.z; acc.alg_y = a0.w;
        acc.att_x = a1.x; acc.att_y = a1.y;
        acc.c_rep = c.x; acc.c_alg = c.y; acc.c_att = c.z;
        if (pred_active) {  // prey<-predator detection of this row was evaluated by detect_rows_kernel
            acc.flee_x = a1.z;
            acc.flee_y = a1.w;
            acc.c_flee = c.w;
        }
    }
    uint32_t fl = 0;
    sbc_agent_step(a, acc, u_social, u_dir, k_next, P, fl, P.BC);
    S.pv[g] = make_float4(a.x, a.y, a.vx, a.vy);
    S.hv[g] = make_float4(a.phi, a.vproj, a.cu, a.su);
    S.force[g] = make_float2(a.fx, a.fy);
    S.timers[g] = make_uint2(a.bin_step, a.stb);
    if (S.flags) S.flags[g] |= (uint8_t)fl;
    if (net_kill) {
        // FishNetKill (agents_dynamics.cpp:780-820) against the net AFTER this step's MoveFishNet (:616-624): the
        // test reads the first two nodes only, so every thread moves its own copy of them (net_move_kernel, which
        // runs behind this kernel, moves the stored net with the same arithmetic)
        const int NP = S.Npred;
        float4 n0 = S.pred_pv[(size_t)rep * NP], n1 = S.pred_pv[(size_t)rep * NP + 1];
        float ph0 = S.pred_phi[(size_t)rep * NP], ph1 = S.pred_phi[(size_t)rep * NP + 1];
        n0.x += n0.z * P.dt; n0.y += n0.w * P.dt;
        sbc_boundary(n0.x, n0.y, n0.z, n0.w, ph0, P, P.BC);
        n1.x += n1.z * P.dt; n1.y += n1.w * P.dt;
        sbc_boundary(n1.x, n1.y, n1.z, n1.w, ph1, P, P.BC);
        if (sbc_net_kills(a.x, a.y, n0, n1, ph0, NP, P)) {  // split_dead (agents_operation.cpp:209-218)
            S.alive[g] = 0;
            S.pv_dead[g] = make_float4(a.x, a.y, a.vx, a.vy);
            S.pv[g] = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), a.vx, a.vy);
            atomicAdd(S.n_dead + rep, 1);
        }
    }
}

}  // namespace

void sbc_launch_update(const DevParams &P, const DevState &S, long long step, const uint32_t *step_dev,
                       int pred_active, cudaStream_t st, bool zero_count) {
    const size_t total = (size_t)S.N * S.R;
    const int net_kill = (pred_active && S.Npred > 1 && P.pred_kill != 0) ? 1 : 0;
    update_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P, S, (uint32_t)step, step_dev, pred_active,
                                                                  zero_count ? S.active_count : nullptr, net_kill);
}
