// ref_wrapper.cpp - TEST INFRASTRUCTURE ONLY ("reference" oracle, builds oracle/_ref/liborc_ref.so).
//
// Thin C interface (oracle/sbc_oracle.h) around the REFERENCE'S OWN translation units, which
// oracle/Makefile compiles in place from /root/reference/src against oracle/shims (no reference
// source is copied into this repository). Everything numerical below is done by reference
// functions: IntCalcPrey/IntCalcPred/InteractionGlobal (agents_interact.cpp), SFM_4Zone
// (social_forces.cpp), ParticleBurstCoast/Boundary/Create*/Move*/…Kill (agents_dynamics.cpp),
// split_dead/GetCenterOfMass (agents_operation.cpp), InitSystem/InitPredator (settings.cpp) and,
// for the Voronoi mode, Step itself (swarmdyn.cpp). This file only moves flat arrays in and out
// and feeds the GSL shim's replay queue with the decisions of the repo's Philox protocol.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "swarmdyn.h"  // reference header: pulls agents.h, agents_interact.h, agents_dynamics.h, ...
#include "sbc_oracle.h"
#include "orc_philox.h"

extern gsl_rng *r;  // the reference's global generator, swarmdyn.cpp:29
int sbc_ref_swarmdyn_main(int argc, char **argv);  // the reference's main(), renamed by -Dmain=

// ---- h5tools.h is declared by the reference but h5tools.cpp is not built (no libhdf5) ----------
static void h5_unavailable(const char *what) {
    std::fprintf(stderr, "oracle/_ref: HDF5 call %s reached but libhdf5 is not available; use -J 0\n", what);
    std::abort();
}
H5::DataSet h5CreateDSet(H5::H5File *, std::vector<hsize_t>, std::string, std::string) { h5_unavailable("h5CreateDSet"); return H5::DataSet(); }
H5::DataSet h5CreateDSetExt(H5::H5File *, std::vector<hsize_t>, std::string, std::string) { h5_unavailable("h5CreateDSetExt"); return H5::DataSet(); }
void h5WriteInt(H5::DataSet *, std::vector<int>, std::vector<hsize_t>) { h5_unavailable("h5WriteInt"); }
void h5WriteDouble(H5::DataSet *, std::vector<double>, std::vector<hsize_t>) { h5_unavailable("h5WriteDouble"); }
void h5readDimension(H5::DataSet *, std::vector<hsize_t> &) { h5_unavailable("h5readDimension"); }
void h5read_extend_dim(H5::DataSet *, std::vector<hsize_t> &) { h5_unavailable("h5read_extend_dim"); }
void h5CreateWriteDset(std::string, std::string, std::string, std::vector<std::vector<double> >, bool) { h5_unavailable("h5CreateWriteDset"); }

struct orc_world {
    params SP;
    sbc_params p;
    int N, Npred;
    uint64_t seed, replicate;
    std::vector<particle> agent, agent_dead;
    std::vector<predator> preds;
    gsl_rng *rng;
    std::vector<uint32_t> geo;
    bool inj_social, inj_dir, inj_k;
    std::vector<double> u_social, u_dir;
    std::vector<uint32_t> k_next;
};

static void fill_params(params &SP, const sbc_params &p, int N, int Npred) {
    SetCoreParameters(&SP);
    SP.N = N;
    SP.Npred = Npred;
    SP.Ndead = 0;
    SP.sizeL = p.sizeL;
    SP.dt = p.dt;
    SP.beta = p.beta;
    SP.alphaTurn = p.alphaTurn;
    SP.soc_strength = p.soc_strength;
    SP.env_strength = p.env_strength;
    SP.prob_social = p.prob_social;
    SP.burst_rate = p.burst_rate;
    SP.rep_range = p.rep_range;
    SP.alg_range = p.alg_range;
    SP.att_range = p.att_range;
    SP.flee_range = p.flee_range;
    SP.pred_time = p.pred_time;
    SP.pred_speed0 = p.pred_speed0;
    SP.kill_range = p.kill_range;
    SP.kill_rate = p.kill_rate;
    SP.noisep = p.noisep;
    SP.sim_time = p.sim_time;
    SP.trans_time = p.trans_time;
    SP.burst_steps = p.burst_steps;
    SP.BC = p.BC;
    SP.N_confu = p.N_confu;
    SP.pred_kill = p.pred_kill;
    SP.pred_move = p.pred_move;
    SP.output_mode = 2;
    SP.out_mean = SP.out_particle = false;
}

static void all_to_alive(orc_world *w) {  // merge_dead keeps ids as indices
    merge_dead(w->agent, w->agent_dead);
}

static void push_geometric(gsl_rng *g, uint32_t k) {
    // the rejection loop agents_dynamics.cpp:257-269 stops at its k-th draw when that draw is
    // <= burst_rate*dt (or unconditionally once steps > max_steps)
    for (uint32_t t = 1; t < k; t++) g->queue.push_back(2.0);
    g->queue.push_back(0.0);
}

static void ref_metric_step(orc_world *w, int64_t s, uint8_t *flags) {
    params *SP = &w->SP;
    gsl_rng *g = w->rng;
    const double dt = SP->dt;
    std::vector<particle> &a = w->agent;
    std::vector<predator> &preds = w->preds;
    const bool have_pred = w->Npred > 0;
    g->replay = true;
    g->queue.clear();
    if (have_pred && s >= SP->pred_time / dt && s - 1 < SP->pred_time / dt) {
        const OrcPhilox4 q = orc_draw(w->seed, w->replicate, ORC_PREDATOR_AGENT_ID, s, 0);
        g->queue.push_back(orc_u32_to_unit(q.v[0]));
        if (preds.size() == 1) {
            CreatePredator(a, SP, preds[0], g);
        } else {
            g->queue.push_back(orc_u32_to_unit(q.v[1]));
            CreateFishNet(a, SP, preds, g);
        }
    }
    if (preds.size() > 1 && s >= SP->pred_time / dt) {
        int since = (int)s - static_cast<int>(SP->pred_time / SP->dt);
        int needed;
        if (SP->pred_speed0 > 0) needed = static_cast<int>(SP->sizeL / 2 / (SP->pred_speed0 * SP->dt));
        else needed = static_cast<int>((SP->sim_time - SP->trans_time) / 4 / SP->dt);
        if (since % needed == 0) {
            const OrcPhilox4 q = orc_draw(w->seed, w->replicate, ORC_PREDATOR_AGENT_ID, s, 1);
            g->queue.push_back(orc_u32_to_unit(q.v[0]));
            g->queue.push_back(orc_u32_to_unit(q.v[1]));
            CreateFishNet(a, SP, preds, g);
        }
    }
    for (size_t i = 0; i < a.size(); i++) a[i].NN.resize(0);
    for (size_t j = 0; j < preds.size(); j++) { preds[j].NNset.clear(); preds[j].NN.resize(0); }
    if (w->p.neighbour_mode == SBC_NEIGHBOUR_VORONOI) InteractionVoronoiF2F(a, SP);  // agents_interact.cpp:26
    else InteractionGlobal(a, SP);                                                  // :144
    if (have_pred && !(s < SP->pred_time / dt)) {
        for (size_t i = 0; i < a.size(); i++)
            for (size_t j = 0; j < preds.size(); j++) {
                const OrcPhilox4 q = orc_draw(w->seed, w->replicate, a[i].id, s,
                                              ORC_SLOT_DETECT0 + (uint32_t)(j / 4));
                g->queue.push_back(orc_u32_to_unit(q.v[j % 4]));
                IntCalcPred(a, (int)i, preds[j], SP);
            }
    }
    for (size_t i = 0; i < a.size(); i++) {
        particle &pa = a[i];
        const unsigned id = pa.id;
        const OrcPhilox4 q = orc_draw(w->seed, w->replicate, id, s, ORC_SLOT_DECISION);
        const double us = w->inj_social ? w->u_social[id] : orc_u32_to_unit(q.v[0]);
        const double ud = w->inj_dir ? w->u_dir[id] : orc_u24_to_unit(q.v[1]);
        const uint32_t kn = w->inj_k ? w->k_next[id] : orc_geometric_k(w->geo, q.v[2]);
        // which draws will ParticleBurstCoast consume? (agents_dynamics.cpp:33,84,259)
        const bool first = (pa.bin_step == SP->burst_steps);
        uint8_t fl = 0;
        g->queue.clear();
        if (first) {
            fl |= 1;
            g->queue.push_back(us);
            if (us <= SP->prob_social) fl |= 2;
            else if (pa.counter_flee > 0) fl |= 64;
            else g->queue.push_back(ud);
        }
        const bool rearm = (pa.steps_till_burst <= 1);
        if (rearm) { fl |= 128; push_geometric(g, kn); }
        const double phi0 = pa.phi, vproj0 = pa.vproj;
        const double fx0 = pa.force[0], fy0 = pa.force[1];
        ParticleBurstCoast(pa, SP, g);
        if (!g->queue.empty()) {
            std::fprintf(stderr, "oracle/_ref: %zu injected draws left over (agent %u step %lld)\n",
                         g->queue.size(), id, (long long)s);
            std::abort();
        }
        (void)phi0; (void)vproj0; (void)fx0; (void)fy0;
        if (flags) flags[id] |= fl;
    }
    if (have_pred && s >= SP->pred_time / dt) {
        if (preds.size() > 1) {
            MoveFishNet(preds, SP);
            if (SP->pred_kill != 0) FishNetKill(a, preds, SP);
        } else {
            const OrcPhilox4 q = orc_draw(w->seed, w->replicate, ORC_PREDATOR_AGENT_ID, s, 2);
            const double u1 = (q.v[0] + 1.0) * (1.0 / 4294967296.0);
            const double u2 = orc_u32_to_unit(q.v[1]);
            g->gauss_queue.clear();
            g->gauss_queue.push_back(std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * M_PI * u2));
            MovePredator(preds[0], a, SP, g);
            g->gauss_queue.clear();
            if (SP->pred_kill > 1) {
                // one `luck` draw per kill candidate in vector order (agents_dynamics.cpp:905)
                for (size_t i = 0; i < a.size(); i++) {
                    const double d = CalcDist(preds[0].x, a[i].x, SP->BC, SP->sizeL);
                    if (d <= 4 * SP->kill_range && d < SP->kill_range) {
                        const OrcPhilox4 qa = orc_draw(w->seed, w->replicate, a[i].id, s, ORC_SLOT_DECISION);
                        g->queue.push_back(orc_u32_to_unit(qa.v[3]));
                    }
                }
            }
            if (SP->pred_kill != 0) PredKill(a, preds[0], SP, g);
        }
    }
    g->queue.clear();
    w->inj_social = w->inj_dir = w->inj_k = false;
}

extern "C" {

const char *orc_kind(void) { return "reference"; }

orc_world *orc_create(const sbc_params *p, int n_agents, int n_pred, uint64_t seed,
                      uint64_t replicate_id) {
    orc_world *w = new orc_world();
    w->p = *p;
    w->N = n_agents;
    w->Npred = n_pred;
    w->seed = seed;
    w->replicate = replicate_id;
    fill_params(w->SP, *p, n_agents, n_pred);
    w->agent.resize(n_agents);
    InitSystem(w->agent, w->SP);  // settings.cpp:168
    for (int i = 0; i < n_agents; i++) {
        w->agent[i].counter_alg = w->agent[i].counter_att = 0;  // InitSystem leaves these unset
        w->agent[i].counter_sor = 0;
    }
    w->preds.resize(n_pred);
    InitPredator(w->preds);  // settings.cpp:147
    w->rng = gsl_rng_alloc(gsl_rng_default);
    w->rng->replay = true;
    w->SP.r = w->rng;
    w->inj_social = w->inj_dir = w->inj_k = false;
    w->geo = orc_geometric_table(p->burst_rate, p->dt);
    return w;
}
void orc_destroy(orc_world *w) {
    gsl_rng_free(w->rng);
    delete w;
}

void orc_set_state(orc_world *w, const double *x, const double *y, const double *vx,
                   const double *vy, const double *phi, const double *vproj, const double *fx,
                   const double *fy, const uint32_t *bin_step, const uint32_t *stb,
                   const uint8_t *alive) {
    all_to_alive(w);
    for (int i = 0; i < w->N; i++) {
        particle &a = w->agent[i];
        a.id = i;
        if (x) a.x[0] = x[i];
        if (y) a.x[1] = y[i];
        if (phi) {
            a.phi = phi[i];
            a.u[0] = cos(a.phi);
            a.u[1] = sin(a.phi);
        }
        if (vproj) a.vproj = vproj[i];
        if (vx) a.v[0] = vx[i]; else if (phi) a.v[0] = a.vproj * a.u[0];
        if (vy) a.v[1] = vy[i]; else if (phi) a.v[1] = a.vproj * a.u[1];
        if (fx) a.force[0] = fx[i];
        if (fy) a.force[1] = fy[i];
        if (bin_step) a.bin_step = bin_step[i];
        if (stb) a.steps_till_burst = stb[i];
        if (alive) a.dead = !alive[i];
    }
    w->SP.Ndead = 0;
    for (int i = 0; i < w->N; i++) w->SP.Ndead += w->agent[i].dead;
    split_dead(w->agent, w->agent_dead, w->preds);
}
void orc_get_state(orc_world *w, double *x, double *y, double *vx, double *vy, double *phi,
                   double *vproj, double *fx, double *fy, uint32_t *bin_step, uint32_t *stb,
                   uint8_t *alive) {
    std::vector<const particle *> by_id(w->N, nullptr);
    for (const particle &a : w->agent) by_id[a.id] = &a;
    for (const particle &a : w->agent_dead) by_id[a.id] = &a;
    for (int i = 0; i < w->N; i++) {
        const particle &a = *by_id[i];
        if (x) x[i] = a.x[0];
        if (y) y[i] = a.x[1];
        if (vx) vx[i] = a.v[0];
        if (vy) vy[i] = a.v[1];
        if (phi) phi[i] = a.phi;
        if (vproj) vproj[i] = a.vproj;
        if (fx) fx[i] = a.force[0];
        if (fy) fy[i] = a.force[1];
        if (bin_step) bin_step[i] = a.bin_step;
        if (stb) stb[i] = a.steps_till_burst;
        if (alive) alive[i] = !a.dead;
    }
}
void orc_set_predators(orc_world *w, const double *x, const double *y, const double *phi, int) {
    for (int j = 0; j < w->Npred; j++) {
        predator &pr = w->preds[j];
        pr.x[0] = x[j];
        pr.x[1] = y[j];
        pr.phi = phi[j];
        pr.u[0] = cos(phi[j]);
        pr.u[1] = sin(phi[j]);
        pr.v[0] = w->SP.pred_speed0 * pr.u[0];
        pr.v[1] = w->SP.pred_speed0 * pr.u[1];
    }
}
void orc_get_predators(orc_world *w, double *out) {
    for (int j = 0; j < w->Npred; j++) {
        std::vector<double> o = w->preds[j].out();  // agents.cpp:20-30
        for (int k = 0; k < 6; k++) out[6 * j + k] = o[k];
    }
}
void orc_get_counts(orc_world *w, int32_t *n_alive, int32_t *n_dead) {
    int d = (int)w->agent_dead.size();
    for (const particle &a : w->agent) d += a.dead;
    if (n_alive) *n_alive = w->N - d;
    if (n_dead) *n_dead = d;
}

int64_t orc_pair_interactions(orc_world *w, int flags, int32_t *c_rep, int32_t *c_alg,
                              int32_t *c_att, double *f_rep, double *f_alg, double *f_att,
                              int64_t *nn_ptr, int32_t *nn_idx, int64_t nn_capacity) {
    std::vector<particle> &a = w->agent;
    std::vector<unsigned> saved(a.size());
    for (size_t i = 0; i < a.size(); i++) {
        particle &pa = a[i];
        pa.force_rep[0] = pa.force_rep[1] = pa.force_alg[0] = pa.force_alg[1] = 0;
        pa.force_att[0] = pa.force_att[1] = 0;
        pa.counter_rep = pa.counter_alg = pa.counter_att = 0;
        pa.NN.resize(0);
        saved[i] = pa.bin_step;
        if (flags == SBC_PAIR_ALL_ROWS) pa.bin_step = w->SP.burst_steps;  // open the gate :178
    }
    if (w->p.neighbour_mode == SBC_NEIGHBOUR_VORONOI) InteractionVoronoiF2F(a, &w->SP);
    else InteractionGlobal(a, &w->SP);
    for (size_t i = 0; i < a.size(); i++) a[i].bin_step = saved[i];
    std::vector<const particle *> by_id(w->N, nullptr);
    for (const particle &pa : a) by_id[pa.id] = &pa;
    int64_t total = 0;
    for (int i = 0; i < w->N; i++) {
        const particle *pa = by_id[i];
        if (c_rep) c_rep[i] = pa ? pa->counter_rep : 0;
        if (c_alg) c_alg[i] = pa ? pa->counter_alg : 0;
        if (c_att) c_att[i] = pa ? pa->counter_att : 0;
        if (f_rep) { f_rep[2 * i] = pa ? pa->force_rep[0] : 0; f_rep[2 * i + 1] = pa ? pa->force_rep[1] : 0; }
        if (f_alg) { f_alg[2 * i] = pa ? pa->force_alg[0] : 0; f_alg[2 * i + 1] = pa ? pa->force_alg[1] : 0; }
        if (f_att) { f_att[2 * i] = pa ? pa->force_att[0] : 0; f_att[2 * i + 1] = pa ? pa->force_att[1] : 0; }
        if (nn_ptr) nn_ptr[i] = total;
        if (pa) {
            std::vector<int> ids;
            for (unsigned pos : pa->NN) ids.push_back((int)a[pos].id);
            std::sort(ids.begin(), ids.end());
            if (nn_idx)
                for (size_t k = 0; k < ids.size(); k++)
                    if (total + (int64_t)k < nn_capacity) nn_idx[total + k] = ids[k];
            total += (int64_t)ids.size();
        }
    }
    if (nn_ptr) nn_ptr[w->N] = total;
    // leave accumulators as a fresh step expects them
    for (particle &pa : a) {
        pa.force_rep[0] = pa.force_rep[1] = pa.force_alg[0] = pa.force_alg[1] = 0;
        pa.force_att[0] = pa.force_att[1] = 0;
        pa.counter_rep = pa.counter_alg = pa.counter_att = 0;
        pa.NN.resize(0);
    }
    return total;
}

void orc_inject_decisions(orc_world *w, const double *u_social, const double *u_dir,
                          const uint32_t *k_next) {
    w->inj_social = u_social != nullptr;
    w->inj_dir = u_dir != nullptr;
    w->inj_k = k_next != nullptr;
    if (u_social) w->u_social.assign(u_social, u_social + w->N);
    if (u_dir) w->u_dir.assign(u_dir, u_dir + w->N);
    if (k_next) w->k_next.assign(k_next, k_next + w->N);
}

void orc_step(orc_world *w, int64_t s_first, int64_t n_steps, uint8_t *branch_flags) {
    for (int64_t s = s_first; s < s_first + n_steps; s++) {
        split_dead(w->agent, w->agent_dead, w->preds);  // swarmdyn.cpp:71
        if (w->agent.size() == 0) break;                // :72-75
        ref_metric_step(w, s, branch_flags);
    }
    split_dead(w->agent, w->agent_dead, w->preds);
}

void orc_philox_decisions(orc_world *w, int64_t step, double *u_social, double *u_dir,
                          uint32_t *k_next) {
    for (int i = 0; i < w->N; i++) {
        const OrcPhilox4 q = orc_draw(w->seed, w->replicate, (uint32_t)i, step, ORC_SLOT_DECISION);
        if (u_social) u_social[i] = orc_u32_to_unit(q.v[0]);
        if (u_dir) u_dir[i] = orc_u24_to_unit(q.v[1]);
        if (k_next) k_next[i] = orc_geometric_k(w->geo, q.v[2]);
    }
}

void orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const OrcPhilox4 q = orc_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int k = 0; k < 4; k++) out[k] = q.v[k];
}

void orc_observables(orc_world *w, double *out) {
    std::vector<particle> &a = w->agent;
    const int n = (int)a.size();
    out[0] = n;
    if (n == 0) { for (int k = 1; k < 8; k++) out[k] = 0; return; }
    double ax[2] = {0, 0}, av[2] = {0, 0}, au[2] = {0, 0}, as = 0, nnd = 0;
    for (int i = 0; i < n; i++) {
        ax[0] += a[i].x[0]; ax[1] += a[i].x[1];
        av[0] += a[i].v[0]; av[1] += a[i].v[1];
        au[0] += a[i].u[0]; au[1] += a[i].u[1];
        as += sqrt(a[i].v[0] * a[i].v[0] + a[i].v[1] * a[i].v[1]);
        double best = -1;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            const double d = CalcDist(a[i].x, a[j].x, w->SP.BC, w->SP.sizeL);
            if (best < 0 || d < best) best = d;
        }
        if (best >= 0) nnd += best;
    }
    au[0] /= n; au[1] /= n;
    out[1] = sqrt(au[0] * au[0] + au[1] * au[1]);
    out[2] = ax[0] / n; out[3] = ax[1] / n;
    out[4] = av[0] / n; out[5] = av[1] / n;
    out[6] = as / n;
    out[7] = nnd / n;
}

/* the reference's own 16-column Out_swarm row (swarmdyn.cpp:237-344) on the alive agents */
void orc_ref_out_swarm(orc_world *w, double *out16) {
    std::vector<double> o = Out_swarm(w->agent, w->SP);
    for (int k = 0; k < 16; k++) out16[k] = o[k];
}

// the reference's own ResetSystem (settings.cpp:196-387) fed with the IC protocol's uniforms through the replay queue.
// The ring shuffle of IC 3-5 is the reference's own (clock-seeded): WHICH agent sits on a ring place differs from the
// protocol's bijection, the set of places does not.
int orc_init_state(orc_world *w, int ic) {
    if (ic == 99) return -1;
    all_to_alive(w);
    InitSystem(w->agent, w->SP);
    gsl_rng *g = w->rng;
    g->replay = true;
    g->queue.clear();
    for (double u : orc_ic_uniform_stream(w->seed, w->replicate, ic, w->N)) g->queue.push_back(u);
    w->SP.IC = ic;
    w->SP.N = w->N;
    ResetSystem(w->agent, &w->SP, false, g);
    const bool dry = g->queue.empty();
    g->queue.clear();
    w->SP.Ndead = 0;
    return dry ? 0 : -2;    // -2: the reference consumed fewer uniforms than the protocol lists
}

// the reference's own rows, for both kinds of library under the same names
void orc_out_swarm(orc_world *w, double *out16) { orc_ref_out_swarm(w, out16); }
void orc_out_swarm_fishnet(orc_world *w, double *out4) {
    std::vector<double> o = Out_swarm_fishNet(w->agent, w->preds, w->SP);
    for (int k = 0; k < 4; k++) out4[k] = o[k];
}

int orc_ref_voronoi_step(orc_world *w, int64_t s_first, int64_t n_steps, uint32_t mt_seed) {
    // the shipped Step (swarmdyn.cpp:143-204): Delaunay drivers, global mt19937 stream
    if (!r) r = gsl_rng_alloc(gsl_rng_default);
    r->replay = false;
    if (mt_seed) gsl_rng_set(r, mt_seed);
    w->SP.r = r;
    for (int64_t s = s_first; s < s_first + n_steps; s++) {
        split_dead(w->agent, w->agent_dead, w->preds);
        if (w->agent.size() == 0) break;
        Step((int)s, w->agent, &w->SP, w->preds);
    }
    w->SP.r = w->rng;
    return 0;
}

double orc_timed_steps(orc_world *w, int64_t s_first, int64_t n_steps) {
    const auto t0 = std::chrono::steady_clock::now();
    orc_step(w, s_first, n_steps, nullptr);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* run the reference's complete main() (text output with -J 0) */
int orc_ref_swarmdyn_main(int argc, char **argv) { return sbc_ref_swarmdyn_main(argc, argv); }

}  // extern "C"
