// sbc_kernels.cuh - host-callable launchers of the sm_100a kernels (definitions in the .cu files).
#pragma once
#include "sbc_device.cuh"

// device-resident state of one handle: R replicates x N agents, index g = r*N + i
// Persistent state of the multi-kernel path: pv 16 + hv 8 + force 8 + tm 4 = 36 B per agent. The
// ensemble kernels additionally keep the heading vector `cs` they turn incrementally, so that a launch can be split
// anywhere bit-exactly; the multi-kernel path derives it from phi and never touches `cs`.
struct DevState {
    int N, R, Npred;
    float4 *pv;        // {x, y, vx, vy}; x = y = NaN for dead agents (their last pv is in pv_dead)
    float4 *pv_nxt;    // multi-kernel path: the update writes the next positions here while the pair pass still reads
                       // pv (start-of-step positions); the host swaps the two after every step
    float4 *pv_dead;   // last {x,y,vx,vy} of killed agents
    float2 *hv;        // {phi, vproj}
    float2 *cs;        // {cos phi, sin phi} as the ensemble kernels carry it (valid when the host says so)
    float2 *force;     // {fx, fy}
    uint32_t *tm;      // burst timers, sbc_tm_pack(bin_step, steps_till_burst)
    uint32_t *tm_nxt;  // multi-kernel path: written by the step, swapped with tm together with pv / pv_nxt - the bulk
                       // update tells the rows (bin_step == burst_steps) by the START-of-step timers while the rows
                       // kernel is already writing theirs
    uint8_t *alive;
    uint32_t *gid;     // global agent id per slot (domain decomposition), nullptr = slot index
    // row accumulators written by the pair pass, consumed by the update
    float *acc;        // [R*N][8] rep.xy alg.xy att.xy flee.xy
    int *cnt;          // [R*N][4] c_rep c_alg c_att c_flee
    // predators
    float4 *pred_pv;   // [R][Npred] {x,y,vx,vy}
    float *pred_phi;   // [R][Npred]
    int *n_dead;       // [R]
    int *pred_spawned; // [R]
    const double *kill_eff; // [N+1] effective kill rate for n sensed prey (PredKill, agents_dynamics.cpp:879-884)
    int *kill_list;         // [R*N] agents the fish net caught during the running step (only with a net), kill_count [1]
    int *kill_count;
    // decisions
    const uint32_t *geo; // inter-burst table, max_steps entries
    double *inj_social;  // [R*N] or nullptr
    float *inj_dir;
    uint32_t *inj_k;
    uint8_t *flags;      // [R*N] branch flags OR-ed over steps (test hook), may be nullptr
    // burst-gated rows: three lists in rotation. The rows of step s are in active_list[s % 3] - written by the update of
    // step s - 1 (the agents it re-armed), or by the compaction kernel after a state change; step s appends the rows of
    // step s + 1 to list (s + 1) % 3 and clears the counter of list (s + 2) % 3, which nobody touches during step s
    int2 *active_list[3];  // [R*N] each: {slot, the slot's cell key at list time (0 without a cell list)}
    int2 *edge_list[3];    // decomposed swarm: the rows next to a slab edge are listed apart (they wait for the ghosts)
    int *active_count;     // [3] (+ 1 pad) of active_list, [4..6] of edge_list
    const uint32_t *cell_key;  // cell key of every slot as of the last cell build (nullptr without a cell list)
    // ensemble kernel in queue mode (swarm_block.cu, EnsQueue): unit counter and slices done per group of tanks
    unsigned *ens_queue;
    int *ens_done;
    // counters
    unsigned long long *stats; // [8] see sbc_get_stats
    // timeline of the multi-kernel step (test / profiling hook, sbc_debug_trace): per step (mod SBC_TRACE_STEPS) and kernel
    // id the wall-clock time the first block started and the last block ended, nullptr = off
    unsigned long long *trace;
};
#define SBC_TRACE_STEPS 1024
enum { SBC_TR_ROWS = 0, SBC_TR_BULK = 1, SBC_TR_RECV = 2, SBC_TR_SEND = 3, SBC_TR_LASTSTART = 4, SBC_TR_EDGE = 6, SBC_TR_KERNELS = 8 };

// per-launch constants of one multi-kernel step (rows kernel + update_bulk_kernel, update.cu)
struct StepArgs {
    uint32_t step;           // step index; with step_dev the offset of this step inside its graph launch
    const uint32_t *step_dev;  // graph-replayed steps: device-resident index of the launch's first step (the last node of
                             // the graph moves it on)
    int n_update;            // slots [0, n_update) can hold agents that are updated (a slab: own_cap; otherwise R * N)
    int net_kill;            // a fish net is present and kills: FishNetKill is evaluated per agent
    int pred_active;         // predator phase: burst-starting rows evaluate IntCalcPred
    int finish;              // rows kernel: 1 = update the rows (a step), 0 = only write acc / cnt (test hook)
    int fresh_cells;         // the cell order was rebuilt after the rows were listed: their cached cell keys are stale
    // decomposed swarm: the neighbours' records of the previous step arrive WHILE this step's kernels run. Rows whose
    // 3 x 3 cells can hold a ghost (cell column <= edge_lo or >= edge_hi) hold before they read candidate positions
    // until ghost_ready[0] shows exchange number step - 1 + ghost_ready[1]; nullptr: nothing to wait for
    const int *ghost_ready;
    int edge_lo, edge_hi;    // edge_hi < 0: not a decomposed swarm, nobody is listed apart
    int edge_blocks;         // rows kernel: the last edge_blocks blocks of the grid take the edge list
    // peer-memory exchange, delivery by the updating thread: an agent listed for a refresh exchange (halo_idx[slot] =
    // its place in the records going left / right, -1 = not listed) stores its new position straight into the
    // neighbours' inboxes xdst[side][parity of the exchange number = step + xshift[0]]; nullptr: nobody delivers
    const int2 *halo_idx;
    void *xdst[2][2];
    const int *xshift;
    float band_lo, band_hi;  // only agents with x < band_lo or x >= band_hi can be listed (saves the halo_idx load)
    int cell_ncx;            // cells per axis (to tell a row's cell column from its key)
    float cell_inv_w;        // cells per unit length (a rebuild step derives the new key of a listed agent itself)
    unsigned long long wait_ns;
};
#ifdef __CUDACC__
// is the cell with this key next to a slab edge (its 3 x 3 block can hold a ghost)?
__device__ __forceinline__ bool sbc_edge_cell(int key, const StepArgs &A) {
    if (A.edge_hi < 0) return false;
    const int cx = key % A.cell_ncx;
    return cx <= A.edge_lo || cx >= A.edge_hi;
}
#endif
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t sbc_step_of(const StepArgs &A) { return A.step_dev ? *A.step_dev + A.step : A.step; }
#endif

// prey<-predator detection (IntCalcPred, /root/reference/src/agents_interact.cpp:252-292) of ONE burst-starting row by
// ONE warp: a lane takes four predators per Philox block; flee sums and counter go to acc[6..7] / cnt[3] of the row.
// Called by the pair kernels right after the row's zone sums (every lane of the warp must call it).
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long sbc_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// every block calls this at entry (begin) and exit: min of the entries, max of the exits
__device__ __forceinline__ void sbc_trace(const DevState &S, uint32_t step, int kid, bool begin) {
    if (!S.trace || threadIdx.x != 0) return;
    unsigned long long *slot = S.trace + ((size_t)(step % SBC_TRACE_STEPS) * SBC_TR_KERNELS + kid) * 2;
    const unsigned long long now = sbc_now_ns();
    if (begin) {
        atomicMin(slot, now);
        // when the LAST block of the kernel started (kernel id + SBC_TR_LASTSTART, entry 1): tells a late dispatch of
        // the grid from a slow kernel
        atomicMax(S.trace + ((size_t)(step % SBC_TRACE_STEPS) * SBC_TR_KERNELS + kid + SBC_TR_LASTSTART) * 2 + 1, now);
    } else {
        atomicMax(slot + 1, now);
    }
}
// agent state in / out of the SoA arrays. `with_cs`: the ensemble kernels' cached heading vector
__device__ __forceinline__ void sbc_load_agent(const DevParams &P, const DevState &S, size_t g, AgentState &a, bool with_cs) {
    const float4 p = S.pv[g];
    const float2 h = S.hv[g];
    const float2 f = S.force[g];
    const uint32_t t = S.tm[g];
    a.x = p.x; a.y = p.y; a.vx = p.z; a.vy = p.w;
    a.phi = h.x; a.vproj = h.y;
    if (with_cs) {
        const float2 c = S.cs[g];
        a.cu = c.x; a.su = c.y;
    }
    a.fx = f.x; a.fy = f.y;
    a.bin_step = sbc_tm_bin(t, P); a.stb = sbc_tm_stb(t, P);
}
// the same past L1 (ld.global.cg), for state another SM wrote during the running launch
__device__ __forceinline__ void sbc_load_agent_cg(const DevParams &P, const DevState &S, size_t g, AgentState &a) {
    const float4 p = __ldcg(S.pv + g);
    const float2 h = __ldcg(S.hv + g);
    const float2 c = __ldcg(S.cs + g);
    const float2 f = __ldcg(S.force + g);
    const uint32_t t = __ldcg(S.tm + g);
    a.x = p.x; a.y = p.y; a.vx = p.z; a.vy = p.w;
    a.phi = h.x; a.vproj = h.y;
    a.cu = c.x; a.su = c.y;
    a.fx = f.x; a.fy = f.y;
    a.bin_step = sbc_tm_bin(t, P); a.stb = sbc_tm_stb(t, P);
}
__device__ __forceinline__ void sbc_store_agent(const DevParams &P, const DevState &S, size_t g, const AgentState &a, bool with_cs) {
    S.pv[g] = make_float4(a.x, a.y, a.vx, a.vy);
    S.hv[g] = make_float2(a.phi, a.vproj);
    if (with_cs) S.cs[g] = make_float2(a.cu, a.su);
    S.force[g] = make_float2(a.fx, a.fy);
    S.tm[g] = sbc_tm_pack(a.bin_step, a.stb, P);
}
// split_dead (agents_operation.cpp:209-218): the agent leaves the active set. NaN positions drop out of every distance
// compare; both position buffers are marked so that the flip of the multi-kernel path cannot resurrect it
__device__ __forceinline__ void sbc_kill_agent(const DevState &S, size_t g, float4 p) {
    S.alive[g] = 0;
    S.pv_dead[g] = p;
    const float qnan = __int_as_float(0x7fc00000);
    const float4 dead = make_float4(qnan, qnan, p.z, p.w);
    S.pv[g] = dead;
    S.pv_nxt[g] = dead;
}

__device__ __forceinline__ void sbc_detect_row_warp(const DevParams &P, const DevState &S, size_t g, float4 pi, uint32_t step,
                                                    int lane, RowAcc &row) {
    const int N = S.N, NP = S.Npred;
    const uint32_t rep = (uint32_t)(g / N);
    const uint32_t i = S.gid ? S.gid[g] : (uint32_t)(g - (size_t)rep * N);
    const float4 *pp = S.pred_pv + (size_t)rep * NP;
    RowAcc acc;
    row_acc_clear(acc);
    for (int jb = lane; 4 * jb < NP; jb += 32) {
        const uint4 q = sbc_draw(P, rep, i, step, SBC_SLOT_DETECT0 + (uint32_t)jb);
        const uint32_t r[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = 4 * jb + u;
            if (j < NP) {
                const float4 pr = pp[j];
                sbc_detect_accumulate(acc, pi.x, pi.y, pr.x, pr.y, (double)r[u] * (1.0 / 4294967296.0), P);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.flee_x += __shfl_xor_sync(0xffffffffu, acc.flee_x, o);
        acc.flee_y += __shfl_xor_sync(0xffffffffu, acc.flee_y, o);
    }
    const int cf = __reduce_add_sync(0xffffffffu, acc.c_flee);
    row.flee_x = acc.flee_x;   // every lane holds the totals
    row.flee_y = acc.flee_y;
    row.c_flee = cf;
    if (lane == 0) {
        S.acc[g * 8 + 6] = acc.flee_x;
        S.acc[g * 8 + 7] = acc.flee_y;
        S.cnt[g * 4 + 3] = cf;
    }
}

// ParticleBurstCoast (agents_dynamics.cpp:150-273) of ONE agent of the multi-kernel path by one thread: state in from
// S.pv (start-of-step position) / hv / force, decisions from the agent's Philox block (or injected), state out to
// S.pv_nxt / hv / force / tm_nxt. `acc` are the row's zone and flee sums (only read when the agent starts a burst).
// Returns true when the agent starts a burst in the NEXT step (it was re-armed and is still alive).
// BCT: the boundary condition when the calling kernel knows it at compile time (the cell-list kernels and the bulk update
// of a periodic box: no wall look-ahead, no reflection code, a fifth fewer instructions fetched and executed), else
// SBC_BC_RUNTIME.
#define SBC_BC_RUNTIME (-99)
template <int BCT = SBC_BC_RUNTIME>
__device__ __forceinline__ bool sbc_update_agent(const DevParams &P, const DevState &S, size_t g, const RowAcc &acc, uint32_t tm,
                                                 float4 p, float2 h, float2 f, uint32_t step, const StepArgs &A) {
    const int BC = (BCT == SBC_BC_RUNTIME) ? P.BC : BCT;
    AgentState a;
    a.x = p.x; a.y = p.y; a.vx = p.z; a.vy = p.w;
    a.phi = h.x; a.vproj = h.y;
    a.fx = f.x; a.fy = f.y;
    a.bin_step = sbc_tm_bin(tm, P); a.stb = sbc_tm_stb(tm, P);
    {
        // heading vector. Where no wall ever edits the velocity (open and periodic domains) v = vproj (cos, sin)(phi)
        // holds at every step boundary (:227-233), so u = v / |v| saves the sincosf; otherwise, and for an agent at
        // rest, it comes from phi as in the reference (:197)
        const float v2 = __fmaf_rn(a.vx, a.vx, a.vy * a.vy);
        if (BC <= 0 && v2 > 1e-30f) {
            const float rinv = rsqrtf(v2);
            a.cu = a.vx * rinv;
            a.su = a.vy * rinv;
        } else {
            sincosf(a.phi, &a.su, &a.cu);
        }
    }
    const bool first = (a.bin_step == P.burst_steps);
    const bool rearm = (a.stb <= 1);
    double u_social = 0.0;
    float u_dir = 0.f;
    uint32_t k_next = 1;
    uint32_t rep = 0;
    if (first || rearm || A.net_kill) {   // (the replicate index costs an integer division: only where it is used)
        rep = (S.R > 1) ? (uint32_t)g / (uint32_t)S.N : 0u;
    }
    if (first || rearm) {
        // Philox counters use the GLOBAL agent id, so a decomposed swarm draws what the whole swarm would
        const uint32_t i = S.gid ? S.gid[g] : (uint32_t)g - rep * (uint32_t)S.N;
        const uint4 q = sbc_draw(P, rep, i, step, SBC_SLOT_DECISION);
        u_social = S.inj_social ? S.inj_social[g] : (double)q.x * (1.0 / 4294967296.0);
        u_dir = S.inj_dir ? S.inj_dir[g] : (float)(q.y >> 8) * (1.0f / 16777216.0f);
        if (rearm) k_next = S.inj_k ? S.inj_k[g] : sbc_geometric_k(S.geo, P.max_steps, q.z, P.inv_log2q);
    }
    uint32_t fl = 0;
    sbc_agent_step(a, acc, u_social, u_dir, k_next, P, fl, BC);
    S.pv_nxt[g] = make_float4(a.x, a.y, a.vx, a.vy);
    S.hv[g] = make_float2(a.phi, a.vproj);
    S.force[g] = make_float2(a.fx, a.fy);
    S.tm_nxt[g] = sbc_tm_pack(a.bin_step, a.stb, P);
    if (S.flags) S.flags[g] |= (uint8_t)fl;
    bool alive = true;
    if (A.net_kill) {
        // FishNetKill (agents_dynamics.cpp:780-820) against the net AFTER this step's MoveFishNet (:616-624): the
        // test reads the first two nodes only, so every thread moves its own copy of them (net_move_kernel, which
        // runs behind the step's kernels, moves the stored net with the same arithmetic)
        const int NP = S.Npred;
        float4 n0 = S.pred_pv[(size_t)rep * NP], n1 = S.pred_pv[(size_t)rep * NP + 1];
        float ph0 = S.pred_phi[(size_t)rep * NP], ph1 = S.pred_phi[(size_t)rep * NP + 1];
        n0.x += n0.z * P.dt; n0.y += n0.w * P.dt;
        sbc_boundary(n0.x, n0.y, n0.z, n0.w, ph0, P, BC);
        n1.x += n1.z * P.dt; n1.y += n1.w * P.dt;
        sbc_boundary(n1.x, n1.y, n1.z, n1.w, ph1, P, BC);
        if (sbc_net_kills(a.x, a.y, n0, n1, ph0, NP, P)) {
            // other rows may still be reading this agent's start-of-step position: the kill itself (split_dead,
            // agents_operation.cpp:209-218) is applied by net_move_kernel behind the step's two kernels
            S.kill_list[atomicAdd(S.kill_count, 1)] = (int)g;
            alive = false;
        }
    }
    if (A.halo_idx && (p.x < A.band_lo || p.x >= A.band_hi)) {
        // decomposed swarm: an agent in an edge band delivers its new position to the neighbour itself; the step's last
        // graph node (dom_publish_kernel) then raises the neighbours' flags
        const int2 hi = A.halo_idx[g];
        if ((hi.x & hi.y) != -1) {
            const int par = ((int)step + A.xshift[0]) & 1;
            // (an agent the net has just caught leaves as dead: NaN position, like its own slot after net_move_kernel)
            const float qnan = __int_as_float(0x7fc00000);
            const float4 npv = alive ? make_float4(a.x, a.y, a.vx, a.vy) : make_float4(qnan, qnan, a.vx, a.vy);
            // 64-byte records behind a header record, position in the first 16 bytes (DomRecord, domain.cu)
            // (selects, not a dynamic index: indexing a kernel parameter array would copy the whole struct to local memory)
            void *dl = par ? A.xdst[0][1] : A.xdst[0][0], *dr = par ? A.xdst[1][1] : A.xdst[1][0];
            if (hi.x >= 0) reinterpret_cast<float4 *>(dl)[(size_t)(1 + hi.x) * 4] = npv;
            if (hi.y >= 0) reinterpret_cast<float4 *>(dr)[(size_t)(1 + hi.y) * 4] = npv;
        }
    }
    return alive && a.bin_step == P.burst_steps;
}

// Decomposed swarm: hold until the ghost slots carry the neighbours' positions of the previous step (the exchange
// runs on another branch of the step's graph; StepArgs::ghost_ready). Called by the rows next to a slab edge right
// before they read candidate positions. Gives up after wait_ns (the exchange itself has put the failure on record).
__device__ __forceinline__ void sbc_wait_ghosts(const StepArgs &A, uint32_t step) {
    const int *r = A.ghost_ready;
    const int need = (int)step - 1 + r[1];
    auto ld_gpu = [](const int *p) {   // device-scope acquire: the word is written by the receive kernel beside this one
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    };
    if (ld_gpu(r) >= need) return;
    const unsigned long long t0 = sbc_now_ns();
    while (ld_gpu(r) < need) {
        if (sbc_now_ns() - t0 > A.wait_ns) break;
        __nanosleep(50);
    }
}

// rows kernel, lane / thread 0 of the group that sweeps row g: the row's own state is fetched BEFORE the sweep (the loads
// overlap the sweep's dependent chain), the row is updated after it and passed on if it re-armed
struct RowState {
    float4 p;
    float2 hv, f;
    uint32_t tm;
    uint8_t al;
};
__device__ __forceinline__ RowState sbc_row_prefetch(const DevState &S, size_t g) {
    RowState r;
    r.al = S.alive[g];
    r.tm = S.tm[g];
    r.p = S.pv[g];
    r.hv = S.hv[g];
    r.f = S.force[g];
    return r;
}
// An agent whose steps_till_burst reads <= 1 at the start of a step is re-armed by that step (agents_dynamics.cpp:
// 184-188, 250-252) and starts a burst in the next one: it can be put on the next step's row list before it is
// updated, so that the latency of the list counter's atomic hides behind the update arithmetic.
__device__ __forceinline__ bool sbc_will_rearm(uint32_t tm, const DevParams &P) { return sbc_tm_stb(tm, P) <= 1u; }
template <int BCT = SBC_BC_RUNTIME>
__device__ __forceinline__ void sbc_finish_row(const DevParams &P, const DevState &S, size_t g, const RowAcc &acc, const RowState &r,
                                               uint32_t step, const StepArgs &A) {
    if (!r.al) return;   // killed after it was listed
    int at = -1, key = 0;
    bool edge = false;
    const int nxt = (int)((step + 1u) % 3u);
    if (sbc_will_rearm(r.tm, P)) {   // rare for a row (it re-armed one step ago): the key's round trip only then
        key = S.cell_key ? (int)S.cell_key[g] : 0;   // (this kernel runs behind the cell build, if there was one)
        edge = sbc_edge_cell(key, A);
        at = atomicAdd(S.active_count + nxt + (edge ? 4 : 0), 1);
    }
    sbc_update_agent<BCT>(P, S, g, acc, r.tm, r.p, r.hv, r.f, step, A);
    // (an agent the net caught this step is dropped by the next rows kernel)
    if (at >= 0) (edge ? S.edge_list : S.active_list)[nxt][at] = make_int2((int)g, key);
}
#endif

// --- host <-> device state conversion (state_io.cu) ------------------------------------------
// staging area in device memory holding the caller's arrays in the ABI's own layout
struct StateStage {
    int f32;        // the caller's real arrays hold floats (sbc_set_state_f32) instead of doubles
    double *d[8];   // x y vx vy phi vproj fx fy
    uint32_t *u[2]; // bin_step, steps_till_burst
    uint8_t *al;
};
enum { SBC_ST_X = 1, SBC_ST_Y = 2, SBC_ST_VX = 4, SBC_ST_VY = 8, SBC_ST_PHI = 16, SBC_ST_VPROJ = 32,
       SBC_ST_FX = 64, SBC_ST_FY = 128, SBC_ST_BIN = 256, SBC_ST_STB = 512, SBC_ST_ALIVE = 1024 };
void sbc_launch_pack_state(const DevParams &P, const DevState &S, const StateStage &st, unsigned mask, int have_prev,
                           cudaStream_t stream);
void sbc_launch_init_cond(const DevParams &P, const DevState &S, int ic, double rep_range, const StateStage &st, cudaStream_t stream);
void sbc_launch_unpack_state(const DevParams &P, const DevState &S, const StateStage &st, unsigned mask, cudaStream_t stream);
void sbc_launch_particles(const DevState &S, double *out_dev, uint8_t *alive_dev, int f32, cudaStream_t stream);

// Launch with the highest stream priority as a LAUNCH ATTRIBUTE: it is recorded in the kernel node of a captured graph,
// so the rows kernel's few blocks are dispatched ahead of the thousands of blocks of the bulk update that is launched
// beside it (a stream's own priority does not survive capture).
#ifdef __CUDACC__
template <class... KArgs, class... Args>
static inline cudaError_t sbc_launch_prio(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static int hi = 1 << 30;
    if (hi == (1 << 30)) {
        int lo = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = hi;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// --- all-pairs three-zone pass -------------------------------------------------------------------
// rows of a fresh state into list / count (update.cu); the steps themselves keep the lists up to date
void sbc_launch_compact_rows(const DevParams &P, const DevState &S, bool clear, bool all_alive, int n_scan, int2 *list, int *count,
                             cudaStream_t st);
// burst-gated rows (bin_step == burst_steps) against all agents of their swarm, pair_allpairs.cu. The rows are
// S.active_list[step % 3]; A.finish: the rows are updated as well (one of the two kernels of a step)
int sbc_pair_rows_blocks(const DevState &S);
void sbc_launch_pair_rows(const DevParams &P, const DevState &S, const StepArgs &A, cudaStream_t st);
// burst-gated rows through a cell list (celllist.cu), single large swarm
struct CellScratch {
    uint32_t *key_in = nullptr, *key_out = nullptr;
    int *idx_in = nullptr, *idx_out = nullptr;   // idx_out[r] = agent id at sorted position r
    int2 *cell_range = nullptr;                   // [ncells] {first, one past last} sorted position of the cell's agents
    int *bbox = nullptr;
    void *geom = nullptr;                         // CellGeom on the device
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0, ncells = 0;
    int ncx = 0;                                  // cells per axis of a periodic box
    int key_bits = 0, max_per_axis = 0;
    bool counting = false;                        // sparse grid: counting-sort build (cnt / start: [ncells + 1])
    int *cnt = nullptr, *start = nullptr;
    float skin = 0.f;                             // cells are att_range + skin wide
};
bool sbc_cells_supported(const DevParams &P, int N, int R);
float sbc_cells_skin(const DevParams &P, float want);
cudaError_t sbc_cell_scratch_alloc(CellScratch &C, const DevParams &P, int N);
void sbc_cell_scratch_free(CellScratch &C);
void sbc_launch_cell_build(const DevParams &P, const DevState &S, CellScratch &C, cudaStream_t st);
int sbc_pair_cells_blocks(const DevParams &P, const DevState &S, const CellScratch &C);
bool sbc_pair_cells_lanes(const DevParams &P, const DevState &S, const CellScratch &C, int pred_active);
void sbc_launch_pair_cells(const DevParams &P, const DevState &S, CellScratch &C, const StepArgs &A, cudaStream_t st);
// dense pass (every alive row), pair_dense.cu: Morton order + tiled all pairs + split combine
struct DenseScratch {
    uint32_t *key_in = nullptr, *key_out = nullptr;
    int *idx_in = nullptr, *idx_out = nullptr;   // idx_out[r] = agent id of sorted row r
    unsigned long long *key64_in = nullptr, *key64_out = nullptr;   // ensembles: (replicate, curve index) keys
    float4 *spv = nullptr;                        // agents in sorted order
    float *part_f = nullptr;                      // [6][R][nsplit][N] partial zone sums
    int *part_i = nullptr;                        // [3][R][nsplit][N] partial zone counts
    int *bbox = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int nsplit = 0, jchunk = 0;
    bool sorted_valid = false;
};
cudaError_t sbc_dense_scratch_alloc(DenseScratch &D, int N, int R);
void sbc_dense_scratch_free(DenseScratch &D);
void sbc_launch_dense_sort(const DevParams &P, const DevState &S, DenseScratch &D, cudaStream_t st);
void sbc_launch_pair_dense(const DevParams &P, const DevState &S, DenseScratch &D, cudaStream_t st);
// test hook: CSR neighbour lists (ascending j) for the same row set
void sbc_launch_nn_fill(const DevParams &P, const DevState &S, int all_rows, const long long *nn_ptr,
                        int *nn_idx, cudaStream_t st);

// --- domain decomposition of one large swarm (domain.cu) --------------------------------------------
struct DomainScratch {
    int *block_counts = nullptr, *block_off = nullptr;  // [nblk][4] per-block category counts / offsets
    int *free_stack = nullptr;                          // free owned slots, popped from the top
    int *ctl = nullptr;  // [0] free_top [1] free_base [2..5] totals (leave L/R, halo L/R) [6] overflow events
                         // [7] migrants taken by the last unpack [8] owned [9] ghosts
                         // [10..13] ghost segment sizes of the last full unpack (halo from a, from b, sent l, sent r)
                         // [14] [15] length of the refresh list towards the left / right neighbour
                         // [17] number of the last exchange whose records are in the ghost slots [18] exchange number minus
                         // step index inside sbc_domain_step [20] / [22] block tickets of the refresh send / receive
                         // [21] a wait for the neighbours' records timed out (sticky until the next sbc_set_state)
    int *halo_list = nullptr;  // [2][max_halo + max_mig] slots whose positions a refresh exchange sends left / right
    int2 *halo_idx = nullptr;  // [slots] the inverse: place of a slot in the left / right list, -1 = not listed
    float halo = 0.f;          // width of the edge band that is mirrored on the neighbour (att_range + skin)
    int nblk = 0;
    int own_cap = 0;         // slots [0, own_cap) are owned / free, [own_cap, N) ghosts
    int max_mig = 0, max_halo = 0;  // record capacities of one buffer: [header | max_mig | max_halo]
    float x_lo = 0.f, x_hi = 0.f;
    unsigned long long timeout_ns = 10000000000ull;  // how long a wait kernel holds for a neighbour (wall clock)
    bool enabled = false;
};
size_t sbc_domain_record_bytes(void);
cudaError_t sbc_domain_scratch_alloc(DomainScratch &D, int own_cap, int max_mig, int max_halo, int n_slots);
void sbc_domain_scratch_free(DomainScratch &D);
void sbc_launch_domain_reset(const DevState &S, DomainScratch &D, cudaStream_t st);
void sbc_launch_domain_pack(const DevParams &P, const DevState &S, DomainScratch &D, void *send_l, void *send_r,
                            cudaStream_t st);
void sbc_launch_domain_unpack(const DevState &S, DomainScratch &D, const void *recv_a, const void *recv_b,
                              const void *sent_l, const void *sent_r, int epoch, cudaStream_t st);
void sbc_launch_domain_refresh_pack(const DevState &S, DomainScratch &D, void *send_l, void *send_r, cudaStream_t st);
void sbc_launch_domain_refresh_unpack(const DevState &S, DomainScratch &D, const void *recv_a0, const void *recv_a1,
                                      const void *recv_b0, const void *recv_b1, int epoch, cudaStream_t st);
void sbc_launch_domain_push(DomainScratch &D, const void *send_l, const void *send_r, void *const dst_l[2], void *const dst_r[2],
                            int *const flag_l[2], int *const flag_r[2], int epoch, cudaStream_t st);
void sbc_launch_domain_wait(DomainScratch &D, int *const flag_a[2], int *const flag_b[2], int epoch, cudaStream_t st);
void sbc_launch_domain_refresh_receive(const DevState &S, DomainScratch &D, const uint32_t *step_dev, uint32_t step_off,
                                       const void *const recv_a[2], const void *const recv_b[2], int *const flag_a[2],
                                       int *const flag_b[2], int *const flag_l[2], int *const flag_r[2], cudaStream_t st);
void sbc_launch_domain_publish(DomainScratch &D, const uint32_t *step_dev, uint32_t step_off, int *const flag_l[2],
                               int *const flag_r[2], cudaStream_t st);
void sbc_launch_count_owned(const DevState &S, DomainScratch &D, cudaStream_t st);

// --- per-agent update of everybody who does not start a burst this step (update.cu) --------------
int sbc_update_bulk_blocks(int n_update);
void sbc_launch_update_bulk(const DevParams &P, const DevState &S, const StepArgs &A, cudaStream_t st);
void sbc_launch_refresh_cs(const DevState &S, cudaStream_t st);
void sbc_launch_step_advance(uint32_t *step_dev, int count, cudaStream_t st);

// --- predator spawn / move / kill (predator.cu) ------------------------------------------------
void sbc_launch_pred_spawn(const DevParams &P, const DevState &S, long long step, int respawn_slot,
                           cudaStream_t st);
// step_back: the device-resident step index has already moved on (the kernel runs behind the step's two kernels)
// decomposed swarm: where every rank's predator area lives (device pointers, own rank included), see predator.cu
struct PredBox {
    unsigned char *peer[16];
    int world, rank;
    unsigned long long timeout_ns;
};
size_t sbc_pred_area_bytes(void);
void sbc_pred_preload(void);
void sbc_cells_preload(void);
void sbc_launch_domain_pred_spawn(const DevParams &P, const DevState &S, const PredBox &B, int own_cap, int *ctl, long long step,
                                  int slot, cudaStream_t st);
void sbc_launch_domain_pred_move_kill(const DevParams &P, const DevState &S, const PredBox &B, int own_cap, int *ctl,
                                      unsigned long long *key, int *cand, const StepArgs &A, cudaStream_t st);
void sbc_launch_pred_move_kill(const DevParams &P, const DevState &S, long long step, const uint32_t *step_dev,
                               cudaStream_t st);

// --- one block per swarm, many steps per launch (swarm_block.cu) -------------------------------
// returns false when the configuration does not fit this path (N > 1024)
// predators (if any): s_pred = first step with s >= pred_time/dt, s_trunc = (int)(pred_time/dt),
// needed = fish-net re-spawn period in steps (0 = none)
bool sbc_launch_swarm_block(const DevParams &P, const DevState &S, long long s_first, long long n_steps,
                            cudaStream_t st, long long s_pred = -1, int s_trunc = 0, int needed = 0,
                            bool force_block = false);

// --- one swarm per thread-block cluster, positions shared through DSMEM (swarm_cluster.cu) ---------
bool sbc_cluster_supported(const DevParams &P, int N, int Npred);

cudaError_t sbc_launch_swarm_cluster(const DevParams &P, const DevState &S, long long s_first, long long n_steps,
                                     cudaStream_t st, long long s_pred = -1, int s_trunc = 0, int needed = 0);

// --- observables (observables.cu) --------------------------------------------------------------
// observables.cu: Out_swarm / Out_swarm_fishNet rows of every replicate (results to host memory; scratch is created on first use)
struct ObsScratch;
struct sbc_params;
#include <string>
int sbc_obs_out_swarm(const DevParams &P, const DevState &S, const sbc_params &hp, ObsScratch **scr, int flags,
                      double *out16_host, double *extra_host, double *out8_host, cudaStream_t st, std::string &err);
int sbc_obs_fishnet(const DevParams &P, const DevState &S, ObsScratch **scr, double *out4_host, cudaStream_t st, std::string &err);
void sbc_obs_free(ObsScratch *o);
