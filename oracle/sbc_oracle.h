/* sbc_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * Flat C interface shared by the two CPU oracles of the burst-coast hot path:
 *   - liborc_port.so  (oracle/sbc_oracle.cpp)  : a double-precision restatement written for this
 *                                               repo, every function citing the reference lines
 *                                               it follows ("port");
 *   - oracle/_ref/liborc_ref.so (oracle/ref_wrapper.cpp + the reference's own .cpp files compiled
 *                                               in place from /root/reference/src against
 *                                               oracle/shims) ("reference").
 * Both export the SAME symbols below, so a test can swap one for the other. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load them;
 * the product (libsbc.so and everything under sde_burst_coast_b200/) never does.
 *
 * Parity status: the reference ships no tests, golden vectors or seeds, so
 * the port is pinned against the compiled reference itself: tests/test_oracle_vs_ref.py demands
 * bit-identical doubles from both libraries on the same injected decisions, and
 * tests/golden/ holds vectors generated from liborc_ref.so (script committed beside them).
 *
 * Neighbour rule: METRIC (= InteractionGlobal, agents_interact.cpp:144-157) unless voronoi=1 is
 * requested from the reference build, which then runs its shipped Delaunay drivers on the
 * brute-force Delaunay shim.
 */
#ifndef SBC_ORACLE_H
#define SBC_ORACLE_H
#include <stdint.h>
#include "../include/sbc.h" /* sbc_params only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

/* which implementation answers: "port" or "reference" */
const char *orc_kind(void);

orc_world *orc_create(const sbc_params *p, int n_agents, int n_pred, uint64_t seed,
                      uint64_t replicate_id);
void orc_destroy(orc_world *w);

/* arrays of length n_agents indexed by agent id; NULL keeps the current value */
void orc_set_state(orc_world *w, const double *x, const double *y, const double *vx,
                   const double *vy, const double *phi, const double *vproj, const double *fx,
                   const double *fy, const uint32_t *bin_step, const uint32_t *steps_till_burst,
                   const uint8_t *alive);
void orc_get_state(orc_world *w, double *x, double *y, double *vx, double *vy, double *phi,
                   double *vproj, double *fx, double *fy, uint32_t *bin_step,
                   uint32_t *steps_till_burst, uint8_t *alive);
void orc_set_predators(orc_world *w, const double *x, const double *y, const double *phi,
                       int spawned);
void orc_get_predators(orc_world *w, double *out /* [Npred][6] */);
void orc_get_counts(orc_world *w, int32_t *n_alive, int32_t *n_dead);

/* Three-zone pair pass on the current state (state is not advanced; accumulators are reset
 * first). flags: SBC_PAIR_ACTIVE_ROWS / SBC_PAIR_ALL_ROWS. nn lists sorted ascending by id.
 * Returns the total NN length (written only if it fits nn_capacity). */
int64_t orc_pair_interactions(orc_world *w, int flags, int32_t *c_rep, int32_t *c_alg,
                              int32_t *c_att, double *f_rep, double *f_alg, double *f_att,
                              int64_t *nn_ptr, int32_t *nn_idx, int64_t nn_capacity);

/* Decision injection for the next orc_step(…, 1): per agent id; NULL = Philox. */
void orc_inject_decisions(orc_world *w, const double *u_social, const double *u_dir,
                          const uint32_t *k_next);

/* Advance n_steps steps starting at step index s_first (split_dead + Step, metric mode).
 * branch_flags (may be NULL, [n_agents], OR-ed over the steps):
 *   bit0 burst start, bit1 social cue, bit2 wall redirect, bit3 vproj<0 hack,
 *   bit4 overshoot clamp, bit5 boundary hit, bit6 flee cue, bit7 re-armed */
void orc_step(orc_world *w, int64_t s_first, int64_t n_steps, uint8_t *branch_flags);

/* The decision draws of (agent, step) from the shared Philox protocol. */
void orc_philox_decisions(orc_world *w, int64_t step, double *u_social, double *u_dir,
                          uint32_t *k_next);

/* raw Philox4x32-10 block (known-answer tests) */
void orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Out_swarm subset (swarmdyn.cpp:237-344), same layout as sbc_observables: out[8]. */
void orc_observables(orc_world *w, double *out);
int orc_init_state(orc_world *w, int ic);                  /* ResetSystem settings.cpp:196-387 on the IC protocol */
void orc_out_swarm(orc_world *w, double *out16);          /* swarmdyn.cpp:237-344 */
void orc_out_swarm_fishnet(orc_world *w, double *out4);   /* swarmdyn.cpp:358-410 */

/* reference build only: run the Step of swarmdyn.cpp:143-204 itself (Voronoi drivers on the
 * Delaunay shim, mt19937 stream) - returns -1 from the port. */
int orc_ref_voronoi_step(orc_world *w, int64_t s_first, int64_t n_steps, uint32_t mt_seed);

/* CPU timing leg (bench.py cpu_baseline): run n_steps with the library's own RNG protocol and
 * return wall seconds of the step loop only (std::chrono::steady_clock, no I/O). */
double orc_timed_steps(orc_world *w, int64_t s_first, int64_t n_steps);

#ifdef __cplusplus
}
#endif
#endif
