// sbc_kernels.cuh - host-callable launchers of the sm_100a kernels (definitions in the .cu files).
#pragma once
#include "sbc_device.cuh"

// device-resident state of one handle: R replicates x N agents, index g = r*N + i
struct DevState {
    int N, R, Npred;
    float4 *pv;        // {x, y, vx, vy}; x = y = NaN for dead agents (their last pv is in pv_dead)
    float4 *pv_dead;   // last {x,y,vx,vy} of killed agents
    float4 *hv;        // {phi, vproj, cos phi, sin phi}
    float2 *force;     // {fx, fy}
    uint2 *timers;     // {bin_step, steps_till_burst}
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
    // decisions
    const uint32_t *geo; // inter-burst table, max_steps entries
    double *inj_social;  // [R*N] or nullptr
    float *inj_dir;
    uint32_t *inj_k;
    uint8_t *flags;      // [R*N] branch flags OR-ed over steps (test hook), may be nullptr
    // scratch of the burst-gated pair pass
    int *active_list;    // [R*N]
    int *active_count;   // [1]
    // counters
    unsigned long long *stats; // [8] see sbc_get_stats
};

// prey<-predator detection (IntCalcPred, /root/reference/src/agents_interact.cpp:252-292) of ONE burst-starting row by
// ONE warp: a lane takes four predators per Philox block; flee sums and counter go to acc[6..7] / cnt[3] of the row.
// Called by the pair kernels right after the row's zone sums (every lane of the warp must call it).
#ifdef __CUDACC__
__device__ __forceinline__ void sbc_detect_row_warp(const DevParams &P, const DevState &S, size_t g, float4 pi, uint32_t step,
                                                    int lane) {
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
    if (lane == 0) {
        S.acc[g * 8 + 6] = acc.flee_x;
        S.acc[g * 8 + 7] = acc.flee_y;
        S.cnt[g * 4 + 3] = cf;
    }
}
#endif

// --- host <-> device state conversion (state_io.cu) ------------------------------------------
// staging area in device memory holding the caller's arrays in the ABI's own layout
struct StateStage {
    double *d[8];   // x y vx vy phi vproj fx fy
    uint32_t *u[2]; // bin_step, steps_till_burst
    uint8_t *al;
};
enum { SBC_ST_X = 1, SBC_ST_Y = 2, SBC_ST_VX = 4, SBC_ST_VY = 8, SBC_ST_PHI = 16, SBC_ST_VPROJ = 32,
       SBC_ST_FX = 64, SBC_ST_FY = 128, SBC_ST_BIN = 256, SBC_ST_STB = 512, SBC_ST_ALIVE = 1024 };
void sbc_launch_pack_state(const DevState &S, const StateStage &st, unsigned mask, int have_prev,
                           cudaStream_t stream);
void sbc_launch_unpack_state(const DevState &S, const StateStage &st, unsigned mask, cudaStream_t stream);
void sbc_launch_particles(const DevState &S, double *out_dev, uint8_t *alive_dev, cudaStream_t stream);

// --- all-pairs three-zone pass -------------------------------------------------------------------
// burst-gated rows (bin_step == burst_steps), pair_allpairs.cu: compaction, then all-pairs rows
void sbc_launch_compact_rows(const DevParams &P, const DevState &S, bool clear, cudaStream_t st, bool all_alive = false,
                             bool count_is_zero = false, uint32_t *step_inc = nullptr);
// detect != 0: the rows also evaluate the prey<-predator detection of step `step` (*step_dev when given)
void sbc_launch_pair_rows(const DevParams &P, const DevState &S, cudaStream_t st, int detect = 0, long long step = 0,
                          const uint32_t *step_dev = nullptr);
// burst-gated rows through a cell list (celllist.cu), single large swarm
struct CellScratch {
    uint32_t *key_in = nullptr, *key_out = nullptr;
    int *idx_in = nullptr, *idx_out = nullptr;   // idx_out[r] = agent id at sorted position r
    int *cell_start = nullptr, *cell_end = nullptr;
    int *bbox = nullptr;
    void *geom = nullptr;                         // CellGeom on the device
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0, ncells = 0;
    int key_bits = 0, max_per_axis = 0;
    float skin = 0.f;                             // cells are att_range + skin wide
};
bool sbc_cells_supported(const DevParams &P, int N, int R);
float sbc_cells_skin(const DevParams &P, float want);
cudaError_t sbc_cell_scratch_alloc(CellScratch &C, const DevParams &P, int N);
void sbc_cell_scratch_free(CellScratch &C);
void sbc_launch_cell_build(const DevParams &P, const DevState &S, CellScratch &C, cudaStream_t st);
void sbc_launch_pair_cells(const DevParams &P, const DevState &S, CellScratch &C, cudaStream_t st, int detect = 0,
                           long long step = 0, const uint32_t *step_dev = nullptr);
// dense pass (every alive row), pair_dense.cu: Morton order + tiled all pairs + split combine
struct DenseScratch {
    uint32_t *key_in = nullptr, *key_out = nullptr;
    int *idx_in = nullptr, *idx_out = nullptr;   // idx_out[r] = agent id of sorted row r
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
                         // [16] number of peer-memory exchanges posted so far [20] block ticket of the fused refresh send
    int *halo_list = nullptr;  // [2][max_halo + max_mig] slots whose positions a refresh exchange sends left / right
    float halo = 0.f;          // width of the edge band that is mirrored on the neighbour (att_range + skin)
    int nblk = 0;
    int own_cap = 0;         // slots [0, own_cap) are owned / free, [own_cap, N) ghosts
    int max_mig = 0, max_halo = 0;  // record capacities of one buffer: [header | max_mig | max_halo]
    float x_lo = 0.f, x_hi = 0.f;
    bool enabled = false;
};
size_t sbc_domain_record_bytes(void);
cudaError_t sbc_domain_scratch_alloc(DomainScratch &D, int own_cap, int max_mig, int max_halo);
void sbc_domain_scratch_free(DomainScratch &D);
void sbc_launch_domain_reset(const DevState &S, DomainScratch &D, cudaStream_t st);
void sbc_launch_domain_pack(const DevParams &P, const DevState &S, DomainScratch &D, void *send_l, void *send_r,
                            cudaStream_t st);
void sbc_launch_domain_unpack(const DevState &S, DomainScratch &D, const void *recv_a, const void *recv_b,
                              const void *sent_l, const void *sent_r, cudaStream_t st);
void sbc_launch_domain_refresh_pack(const DevState &S, DomainScratch &D, void *send_l, void *send_r, cudaStream_t st);
void sbc_launch_domain_refresh_unpack(const DevState &S, DomainScratch &D, const void *recv_a0, const void *recv_a1,
                                      const void *recv_b0, const void *recv_b1, cudaStream_t st);
void sbc_launch_domain_push(DomainScratch &D, const void *send_l, const void *send_r, void *const dst_l[2], void *const dst_r[2],
                            int *const flag_l[2], int *const flag_r[2], cudaStream_t st);
void sbc_launch_domain_wait(DomainScratch &D, int *const flag_a[2], int *const flag_b[2], cudaStream_t st);
void sbc_launch_domain_refresh_fused(const DevState &S, DomainScratch &D, void *const dst_l[2], void *const dst_r[2],
                                     int *const flag_l[2], int *const flag_r[2], const void *const recv_a[2],
                                     const void *const recv_b[2], int *const flag_a[2], int *const flag_b[2], cudaStream_t st);
void sbc_launch_count_owned(const DevState &S, DomainScratch &D, cudaStream_t st);

// --- per-agent update incl. prey<-predator detection (update.cu) -------------------------------
// pred_active: the rows' flee sums are valid (the pair kernel ran the detection); zero_count: reset the active-row counter for the next
// step's compaction once nobody reads it any more; the fish-net kill is evaluated here when a net is present
void sbc_launch_update(const DevParams &P, const DevState &S, long long step, const uint32_t *step_dev,
                       int pred_active, cudaStream_t st, bool zero_count = false);

// --- predator spawn / move / kill (predator.cu) ------------------------------------------------
void sbc_launch_pred_spawn(const DevParams &P, const DevState &S, long long step, int respawn_slot,
                           cudaStream_t st);
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
void sbc_launch_observables(const DevParams &P, const DevState &S, double *out_dev, cudaStream_t st);
