/* sbc.h - C ABI of the B200-native burst-coast hot path (libsbc.so).
 *
 * Drop-in boundary: the reference has no FFI layer; its hot path is the
 * single call `Step(s, agent, &SysPara, preds)` at /root/reference/src/swarmdyn.cpp:76, bracketed
 * by `split_dead` (:71) and `Output` (:85). A host `main` (ours: sde_burst_coast_b200/host/
 * swarmdyn_main.cpp; or the reference's own, patched as shown in INTEGRATION.md) keeps argv
 * parsing, initial conditions and file output, and replaces that bracket by the calls below.
 *
 * Conventions: every function returns 0 on success or a negative sbc_status; the message is
 * available from sbc_last_error(). No C++ exception crosses this boundary. All pointers are
 * HOST pointers unless the name ends in `_dev`. Doubles at the ABI (the reference's layout),
 * FP32 structure-of-arrays inside. A handle owns its device memory and is bound to one device
 * and one stream; it is not thread-safe, different handles may be used from different threads.
 *
 * There is NO CPU fallback: every entry point that computes launches CUDA kernels and fails
 * with SBC_ERR_CUDA when no device is usable.
 */
#ifndef SBC_H
#define SBC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBC_ABI_VERSION 1

typedef enum sbc_status {
    SBC_OK = 0,
    SBC_ERR_INVALID = -1, /* bad argument / unsupported parameter combination */
    SBC_ERR_CUDA = -2,    /* CUDA runtime error, no device, or kernel fault */
    SBC_ERR_STATE = -3,   /* call sequence error (e.g. step before set_state) */
    SBC_ERR_NOMEM = -4
} sbc_status;

/* neighbour rule: which candidate set feeds the three-zone filter */
enum { SBC_NEIGHBOUR_METRIC = 0 /* = InteractionGlobal, agents_interact.cpp:144-157 */,
       SBC_NEIGHBOUR_VORONOI = 1 /* = InteractionVoronoiF2F :26-77 / F2FP :79-141 (the shipped default): Delaunay
                                    neighbours, swarms of <= 1024 agents. Periodic boxes: nearest-image
                                    triangulation (what GetCopies4PeriodicBC, agents_operation.cpp:240-293,
                                    means to build); predator phase: predators are vertices, every edge
                                    interacts both ways (the reference: one arbitrary way) */ };
/* pair-search path */
enum { SBC_PATH_AUTO = 0, SBC_PATH_ALLPAIRS = 1, SBC_PATH_CELLLIST = 2 };

/* POD mirror of the hot fields of `params` (/root/reference/src/agents.h:83-148). */
typedef struct sbc_params {
    double sizeL;        /* -L  box length (BC 0..4) or tank radius (BC 5,6)            agents.h:90  */
    double dt;           /* -d                                                           agents.h:112 */
    double beta;         /* -b  friction along heading                                   agents.h:143 */
    double alphaTurn;    /* -Y                                                           agents.h:104 */
    double soc_strength; /* -H                                                           agents.h:94  */
    double env_strength; /* -R                                                           agents.h:95  */
    double prob_social;  /* -Q                                                           agents.h:98  */
    double burst_rate;   /* -A                                                           agents.h:96  */
    double rep_range;    /* -h                                                           agents.h:100 */
    double alg_range;    /* -a                                                           agents.h:101 */
    double att_range;    /* -r                                                           agents.h:102 */
    double flee_range;   /* -f                                                           agents.h:105 */
    double pred_time;    /* -e  predator appears at the first s with s >= pred_time/dt   agents.h:107 */
    double pred_speed0;  /* -S                                                           agents.h:108 */
    double kill_range;   /* -G                                                           agents.h:133 */
    double kill_rate;    /* -O                                                           agents.h:132 */
    double noisep;       /* sqrt(2 dt Dphi), only for pred_move==0                       agents.h:92  */
    double sim_time;     /* -t  (fish-net respawn period when pred_speed0==0)            agents.h:110 */
    double trans_time;   /* -T                                                           agents.h:120 */
    uint32_t burst_steps; /* int(burst_duration/dt)                                      agents.h:142 */
    int32_t BC;          /* -B  -1 open, 0 periodic, 1..4 boxes, 5 elastic / 6 inelastic circle */
    int32_t N_confu;     /* -c                                                           agents.h:116 */
    int32_t pred_kill;   /* -x  0..4                                                     agents.h:117 */
    int32_t pred_move;   /* -X  0 random walk, 1 chase closest                           agents.h:118 */
    int32_t neighbour_mode; /* SBC_NEIGHBOUR_* */
    int32_t path;           /* SBC_PATH_* */
    int32_t reserved;
} sbc_params;

typedef struct sbc_handle sbc_handle;

/* ---- life cycle -------------------------------------------------------------------------- */

/* Create a simulation of `n_replicates` independent swarms of `n_agents` prey and `n_pred`
 * predators (0, 1 = single predator, >1 = fish net) on CUDA device `device`.
 * Replaces the allocation of `agent`, `preds`, `SysPara` at swarmdyn.cpp:38-57.
 * `seed` keys the Philox4x32-10 streams; `replicate_offset` is the global id of this handle's
 * first replicate so that sharded ensembles draw the same numbers for any GPU count. */
int sbc_create(const sbc_params *p, int n_agents, int n_pred, int n_replicates,
               uint64_t seed, uint64_t replicate_offset, int device, sbc_handle **out);
int sbc_destroy(sbc_handle *h);
const char *sbc_last_error(const sbc_handle *h); /* h may be NULL: last create error */
int sbc_abi_version(void);

/* Use an existing CUDA stream (a cudaStream_t passed as void*) instead of the handle's own. */
int sbc_set_stream(sbc_handle *h, void *cuda_stream);
int sbc_synchronize(sbc_handle *h);

/* ---- state in / out (arrays are [n_replicates][n_agents], replicate-major) ----------------- */

/* Initial conditions generated ON THE DEVICE for every replicate: ResetSystem (settings.cpp:196-387) cases IC = 0..5 and
 * the "else" branch (uniform in the box), followed by the reference's Boundary call (:384-386); unit speed, no force,
 * burst timers 0 (InitSystem :168-193). The uniforms the reference pops from its clock-seeded GSL stream come from the
 * handle's Philox key instead (counter: loop index, step 0xFFFFFFFF, slot 16, global replicate id), so a replicate's
 * state does not depend on how the ensemble is split over handles or GPUs. IC 99 (a file) stays on the host. */
int sbc_init_state(sbc_handle *h, int ic);

/* Upload the agents' state; replaces InitSystem/ResetSystem output (settings.cpp:168-387) as the
 * source of x, v, phi, vproj, force, bin_step, steps_till_burst. Any pointer may be NULL to
 * keep the reference's InitSystem default (x,v from phi; vproj=1; force=0; counters=0). */
int sbc_set_state(sbc_handle *h, const double *x, const double *y, const double *vx,
                  const double *vy, const double *phi, const double *vproj, const double *fx,
                  const double *fy, const uint32_t *bin_step, const uint32_t *steps_till_burst,
                  const uint8_t *alive);
int sbc_get_state(sbc_handle *h, double *x, double *y, double *vx, double *vy, double *phi,
                  double *vproj, double *fx, double *fy, uint32_t *bin_step,
                  uint32_t *steps_till_burst, uint8_t *alive);

/* The same with FLOAT arrays - what the device keeps anyway: for callers whose own data is single precision and for
 * host links that are short of bandwidth (half the bytes of the reference layout). Same NULL rules. */
int sbc_set_state_f32(sbc_handle *h, const float *x, const float *y, const float *vx, const float *vy,
                      const float *phi, const float *vproj, const float *fx, const float *fy,
                      const uint32_t *bin_step, const uint32_t *steps_till_burst, const uint8_t *alive);

/* Rows in `particle::out()` layout (agents.cpp:4-17): x0 x1 v0 v1 fitness(=0) force0 force1,
 * indexed by agent id; dead agents keep all-zero rows as in /part (swarmdyn.cpp:561-600). */
int sbc_get_particles(sbc_handle *h, double *out /* [R][N][7] */, uint8_t *alive /* [R][N] or NULL */);
int sbc_get_particles_f32(sbc_handle *h, float *out /* [R][N][7] */, uint8_t *alive /* [R][N] or NULL */);
/* Rows in `predator::out()` layout (agents.cpp:20-30): x0 x1 v0 v1 0 0. */
int sbc_get_predators(sbc_handle *h, double *out /* [R][Npred][6] */);
int sbc_set_predators(sbc_handle *h, const double *x, const double *y, const double *phi,
                      int spawned);
int sbc_get_counts(sbc_handle *h, int32_t *n_alive /* [R] */, int32_t *n_dead /* [R] */);

/* ---- the hot path ------------------------------------------------------------------------- */

/* Advance all replicates by `n_steps` consecutive steps s_first, s_first+1, ...; each step is
 * split_dead + Step (swarmdyn.cpp:71,76,143-204) in metric neighbour mode. */
int sbc_step(sbc_handle *h, int64_t s_first, int64_t n_steps);

/* ---- observables on the device (Out_swarm swarmdyn.cpp:237-344, Out_swarm_fishNet :358-410) ---- */
/* out[R][8]: N_alive, pol_order, avg_x0, avg_x1, avg_v0, avg_v1, avg_speed, NND(true nearest) */
int sbc_observables(sbc_handle *h, double *out);

/* out[R][16] = the reference's /swarm row of every replicate, from the state as it stands:
 *   N, pol_order, L_norm, avg_x0, avg_x1, avg_v0, avg_v1, avg_speed, avg_vsquare, elongation, aspect_ratio,
 *   a_com_maxIID, Area_ConHull, NND, IID, ND                                  (swarmdyn.cpp:324-343)
 * with the reference's quirks (NND skips the agent listed just before i and is capped at N*alg_range :273-275, IID is
 * the sum over j > i divided by N(N-1) :296, ND = NaN :270). extra[R][4] (may be NULL): mean distance to the TRUE
 * nearest neighbour, mean distance over unordered pairs, the largest inter-individual distance, hull vertex count.
 * Selections (nearest neighbour, most distant pair, extreme projections) are made in FP64 in the reference's operation
 * order, so they pick the same agents; sums differ from a sequential double sum by summation order only.
 * SBC_OBS_NO_PAIRS skips the O(N^2) pass: NND, IID, aspect_ratio, a_com_maxIID and extra[0..2] are NaN. */
#define SBC_OBS_NO_PAIRS 1
int sbc_out_swarm(sbc_handle *h, int flags, double *out, double *extra);

/* out[R][4] = the /swarm_fishNet row: NfrontClose, distNet, Nfront, Ndead (swarmdyn.cpp:358-410); needs n_pred >= 1 */
int sbc_out_swarm_fishnet(sbc_handle *h, double *out);

/* ---- test hooks (parity tests call these through the ABI) --------------------------------- */

#define SBC_PAIR_ACTIVE_ROWS 0 /* rows with bin_step == burst_steps (IntCalcPrey gate, :178) */
#define SBC_PAIR_ALL_ROWS 1    /* every alive row (state at s=1) */
#define SBC_PAIR_RESORT 2      /* sbc_bench_pair_pass: rebuild the Morton order inside every timed pass */

/* Run the three-zone pair pass on the current state without advancing it. Outputs per agent
 * ([R][N]; any may be NULL): zone counters, accumulated force_rep/alg/att (x,y interleaved,
 * [R][N][2]) as IntCalcPrey leaves them (agents_interact.cpp:211-225).
 * `nn_ptr` [R*N+1] / `nn_idx` [nn_capacity]: CSR neighbour lists sorted by j (the reference's
 * NN holds the same set in visiting order); pass NULL to skip. Returns SBC_ERR_NOMEM if
 * nn_capacity is too small (nn_ptr[R*N] then holds the needed size). */
int sbc_pair_interactions(sbc_handle *h, int flags, int32_t *c_rep, int32_t *c_alg,
                          int32_t *c_att, double *f_rep, double *f_alg, double *f_att,
                          int64_t *nn_ptr, int32_t *nn_idx, int64_t nn_capacity);

/* Decision-level injection for the NEXT sbc_step call of n_steps == 1:
 * u_social[i] (compared with prob_social, agents_dynamics.cpp:33-35), u_dir[i] (random cue
 * direction 2*pi*u, :84) and k_next[i] (geometric inter-burst draw, :257-269) replace the
 * Philox draws of every agent i; NULL keeps Philox for that quantity. */
int sbc_inject_decisions(sbc_handle *h, const double *u_social, const double *u_dir,
                         const uint32_t *k_next);

/* The exact uniform draws of (replicate, agent, step): u_social, u_dir as doubles and the
 * geometric draw k; lets the CPU oracle replay the device's own Philox stream. */
int sbc_philox_decisions(sbc_handle *h, int64_t step, double *u_social, double *u_dir,
                         uint32_t *k_next);

/* One raw Philox4x32-10 block from the library's host implementation (known-answer tests). */
int sbc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Branch flags of ParticleBurstCoast OR-ed over the steps since the last read, per agent:
 * bit0 burst start, bit1 social cue, bit2 wall redirect, bit3 vproj<0 hack, bit4 overshoot clamp,
 * bit5 boundary hit, bit6 flee cue, bit7 re-armed. enable!=0 starts collecting; out!=NULL reads
 * ([R][N]) and clears. */
int sbc_debug_flags(sbc_handle *h, int enable, uint8_t *out);
/* on = 1: route swarms of <= 1024 agents through the multi-kernel path (pair pass + update kernels)
 * instead of the persistent one-swarm-per-block / per-cluster kernels; on = 2: keep swarms of 257..1024
 * agents on the one-block kernel instead of the thread-block-cluster kernel; 0: automatic. Lets the
 * tests check every schedule on one state. */
int sbc_debug_force_multikernel(sbc_handle *h, int on);

/* Timeline of the multi-kernel step (profiling hook): while enabled, every step records for each of its kernels
 * (0 rows kernel, 1 bulk update, 2 exchange receive, 3 exchange send) the wall-clock nanoseconds (%globaltimer) at
 * which its first block started and its last block ended. out (may be NULL): uint64 [1024][8][2] = {start, end} by
 * (step mod 1024, kernel); exchanges are indexed by their running count instead of the step. Untouched entries read
 * {UINT64_MAX, 0}. enable != 0 (re)starts the recording, 0 stops it. */
int sbc_debug_trace(sbc_handle *h, int enable, uint64_t *out);

/* Kernel statistics since creation: [0] pair-kernel launches, [1] update launches,
 * [2] ensemble launches, [3] directed pairs evaluated, [4] ambiguous (FP64 re-checked) pairs,
 * [5] agent-steps, [6] cell-list builds, [7] other launches. */
int sbc_get_stats(sbc_handle *h, uint64_t out[8]);

/* ---- spatial domain decomposition of ONE large periodic swarm over several handles -----------------
 * (one handle per GPU; SURVEY.md section 8e, BASELINE config C5). The reference is a single process and
 * has no counterpart; Step()'s semantics (swarmdyn.cpp:143-204) are kept: interactions see positions at
 * the start of the step, every agent is updated once, decisions depend on (seed, GLOBAL id, step).
 * A handle's n_agents is its slot capacity: slots [0, own_capacity) hold owned agents or free slots,
 * the rest hold ghosts (copies of the neighbours' agents within att_range + skin of a slab edge).
 * Buffers are DEVICE pointers of sbc_domain_buffer_bytes(max_migrants, max_halo) bytes:
 * [header | migrant records | halo records], 64-byte records, fixed size - no host round trip sizes a
 * message. ONE exchange per step: after sbc_set_state and after every sbc_step(h, s, 1) the host does
 * sbc_domain_pack -> exchange with the two ring neighbours -> sbc_domain_unpack
 * (sde_burst_coast_b200/domain.py; transport = NCCL send/recv over NVLink through torch.distributed).
 * The exchange is FULL (migration, new ghost sets, new cell order in the next step) every period + 1
 * steps and a position REFRESH of the ghosts in place otherwise - the library picks the kind, the same on
 * every rank; see sbc_domain_rebuild_period. */
int sbc_set_global_ids(sbc_handle *h, const uint32_t *gid /* [n_agents] */);
int sbc_get_global_ids(sbc_handle *h, uint32_t *gid /* [n_agents]; 0xFFFFFFFF for empty ghost slots */);
int sbc_domain_config(sbc_handle *h, int rank, int world, int own_capacity, int max_migrants, int max_halo);
size_t sbc_domain_buffer_bytes(int max_migrants, int max_halo);
/* classify the owned agents (left the slab / inside a halo band), write both buffers, free leavers' slots */
int sbc_domain_pack(sbc_handle *h, void *send_left_dev, void *send_right_dev);
/* recv_a / recv_b: what the left / right neighbour sent to this rank; sent_*: this rank's own send buffers
 * (its outgoing migrants become its ghosts) */
int sbc_domain_unpack(sbc_handle *h, const void *recv_a_dev, const void *recv_b_dev, const void *sent_left_dev,
                      const void *sent_right_dev);
int sbc_domain_counts(sbc_handle *h, int32_t *n_owned, int32_t *n_ghost, int32_t *overflow_events);
/* period < 0: returns how many refresh exchanges this rank's own agents allow between two full ones (the
 * Verlet-skin bound of the cell list, from the largest speed in its state). period >= 0: sets the number
 * all ranks agreed on (the minimum of the above; 0 = every exchange is full, the default) and returns the
 * value in effect. Call after sbc_set_state, before the first sbc_domain_pack. */
int sbc_domain_rebuild_period(sbc_handle *h, int period);

/* Exchange through PEER MEMORY instead of a message library: every rank owns an inbox in device memory
 * (two buffers per side, by parity of the exchange count, plus arrival flags); a rank's push kernel stores
 * the used part of its send buffers directly into its neighbours' inboxes over NVLink and then publishes
 * the exchange count in their flags; the receiver's stream holds in a one-thread kernel until both flags
 * show it. No host synchronisation and no collective call per step.
 *   sbc_domain_inbox      : allocate the inbox (once); returns its device address, size and - if asked - the
 *                           64-byte CUDA IPC handle another process opens with sbc_domain_open_peer
 *   sbc_domain_attach_peers: the inboxes of the left / right ring neighbour, as device pointers valid here
 *   sbc_domain_post       : pack + push to both neighbours;  sbc_domain_complete : wait + unpack
 *   sbc_domain_step       : n_steps x { sbc_step(h, s, 1); post; complete }
 * Every rank must make the same sequence of post / complete calls. */
int sbc_domain_inbox(sbc_handle *h, void **base_dev, size_t *bytes, void *ipc_handle_64 /* or NULL */);
int sbc_domain_open_peer(sbc_handle *h, const void *ipc_handle_64, void **mapped_dev);
int sbc_domain_attach_peers(sbc_handle *h, void *left_inbox_dev, void *right_inbox_dev);
/* Predators of a decomposed swarm (n_pred >= 1 at sbc_create): the predator is replicated on every rank; the closest prey
 * of the chase, the sensed-prey count / weight sum of PredKill and the centre of mass at spawn are all-gathered as one
 * 64-byte record per rank through a predator area inside every inbox, so each handle needs the inbox of EVERY rank
 * (device pointers in rank order, entry [rank] = its own; world <= 16). Replaces the reference's serial loops
 * agents_dynamics.cpp:580-597, :861-877 and agents_operation.cpp:23-127 for a swarm cut over GPUs. */
int sbc_domain_attach_world(sbc_handle *h, void *const *inboxes_dev, int n);
int sbc_domain_post(sbc_handle *h);
int sbc_domain_complete(sbc_handle *h);
int sbc_domain_step(sbc_handle *h, int64_t s_first, int64_t n_steps);
/* How long a wait kernel holds for a neighbour's records (wall clock, default 10 s; the environment variable
 * SBC_DOMAIN_TIMEOUT_S overrides the default at sbc_domain_config). A wait that runs out does NOT apply the stale
 * inbox: the handle is marked failed and sbc_domain_step / sbc_domain_complete return SBC_ERR_STATE. The same
 * status reports record-buffer overflows and header mismatches between neighbours. */
int sbc_domain_set_timeout(sbc_handle *h, double seconds);

/* ---- micro-benchmark hook: time `iters` launches of the dense pair pass with CUDA events on
 * the handle's stream; returns mean milliseconds per launch in *ms. */
int sbc_bench_pair_pass(sbc_handle *h, int flags, int iters, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* SBC_H */
