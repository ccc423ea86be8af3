// swarmdyn_main.cpp - host executable `swarmdyn` of the B200 build: the reference's command line,
// parameter derivations, initial conditions and output files, with the per-timestep hot path
// (split_dead + Step, /root/reference/src/swarmdyn.cpp:71,76) replaced by calls into libsbc.so
// (include/sbc.h). Written from scratch; reference lines are cited where behaviour is mirrored.
//
//   ./swarmdyn -L 24.5 -N 8 ... (the 32 two-token flags of src/input_output.cpp:120-151; a missing
//   flag parses as 0 / "0" exactly as getCmdOption :108-114 does)
// New optional flags (the two letters the reference leaves free, input_output.cpp:119):
//   -Z <seed>        64-bit seed of the Philox streams and of the initial conditions (the reference
//                    seeds from the wall clock, swarmdyn.cpp:126-141; 0 or absent = clock)
//   -y <replicates>  run an ensemble of independent swarms in one process (default 1); replicate r>0
//                    writes its files with the suffix _<r> after fileID
//   -V 1             Voronoi neighbour rule (InteractionVoronoiF2F / F2FP, agents_interact.cpp:26-141, the shipped
//                    default of the reference): swarms of up to 1024 agents (DESIGN.md section 2 for the periodic
//                    box and the predator phase)
// Default neighbour rule: METRIC (InteractionGlobal, agents_interact.cpp:144-157) - see DESIGN.md section 2.
// Output: -J 0 text files exactly as the reference writes them; -J 1 writes out_<fileID>.h5 with the
// reference's datasets (/part, /pred, /swarm, /swarm_fishNet; under /<fileID>/ when fileID != "xx") through
// the dependency-free writer h5mini.hpp (contiguous float64 datasets; no libhdf5 in this image). The
// reference's "extend" mode (append a run to an existing file, h5tools.cpp:174-178) is not reproduced:
// every run writes a fresh file.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "../../include/sbc.h"
#include "h5mini.hpp"

namespace {

struct HostParams {  // the fields of `params` (src/agents.h:83-148) the host side needs
    double sizeL, Dphi, beta, dt, sim_time, trans_time, rep_range, alg_range, att_range;
    double soc_strength, burst_rate, env_strength, output, flee_range, pred_time, pred_speed0;
    double kill_range, kill_rate, prob_social, burst_duration, alphaTurn, noisep;
    int N, Npred, BC, IC, N_confu, output_mode, pred_kill, pred_move, out_h5;
    int sim_steps, step_output, total_outstep, total_outstep_pred, burst_steps;
    bool out_extend, out_mean, out_particle;
    std::string location, fileID;
    uint64_t seed;
    int replicates;
    int voronoi;
};

const char *cmd_option(char **begin, char **end, const std::string &opt) {  // input_output.cpp:108-114
    char **it = std::find(begin, end, opt);
    if (it != end && ++it != end) return *it;
    return "0";
}

void parse_parameters(int argc, char **argv, HostParams &p) {  // input_output.cpp:116-158
    char **b = argv, **e = argv + argc;
    p.sizeL = atof(cmd_option(b, e, "-L"));
    p.location = cmd_option(b, e, "-l");
    p.N = atoi(cmd_option(b, e, "-N"));
    p.Npred = atoi(cmd_option(b, e, "-n"));
    p.flee_range = atof(cmd_option(b, e, "-f"));
    p.pred_time = atof(cmd_option(b, e, "-e"));
    p.pred_speed0 = atof(cmd_option(b, e, "-S"));
    p.Dphi = atof(cmd_option(b, e, "-D"));
    p.beta = atof(cmd_option(b, e, "-b"));
    p.dt = atof(cmd_option(b, e, "-d"));
    p.sim_time = atof(cmd_option(b, e, "-t"));
    p.trans_time = atof(cmd_option(b, e, "-T"));
    p.rep_range = atof(cmd_option(b, e, "-h"));
    p.alg_range = atof(cmd_option(b, e, "-a"));
    p.att_range = atof(cmd_option(b, e, "-r"));
    p.soc_strength = atof(cmd_option(b, e, "-H"));
    p.burst_rate = atof(cmd_option(b, e, "-A"));
    p.env_strength = atof(cmd_option(b, e, "-R"));
    p.output = atof(cmd_option(b, e, "-o"));
    p.BC = atoi(cmd_option(b, e, "-B"));
    p.IC = atoi(cmd_option(b, e, "-I"));
    p.N_confu = atoi(cmd_option(b, e, "-c"));
    p.output_mode = atoi(cmd_option(b, e, "-m"));
    p.kill_range = atof(cmd_option(b, e, "-G"));
    p.pred_kill = atoi(cmd_option(b, e, "-x"));
    p.pred_move = atoi(cmd_option(b, e, "-X"));
    p.fileID = cmd_option(b, e, "-E");
    p.out_h5 = atoi(cmd_option(b, e, "-J"));
    p.prob_social = atof(cmd_option(b, e, "-Q"));
    p.burst_duration = atof(cmd_option(b, e, "-u"));
    p.alphaTurn = atof(cmd_option(b, e, "-Y"));
    p.kill_rate = atof(cmd_option(b, e, "-O"));
    p.step_output = (p.output > p.dt) ? (int)(p.output / p.dt) : 1;
    p.sim_steps = (int)(p.sim_time / p.dt);
    p.seed = strtoull(cmd_option(b, e, "-Z"), nullptr, 10);
    p.replicates = atoi(cmd_option(b, e, "-y"));
    if (p.replicates < 1) p.replicates = 1;
    p.voronoi = atoi(cmd_option(b, e, "-V"));
}

void init_system_parameters(HostParams &p) {  // settings.cpp:79-145
    p.noisep = std::sqrt(2 * p.dt * p.Dphi);
    if (p.output < p.dt) p.output = p.dt;
    p.step_output = (int)(p.output / p.dt);
    if (p.trans_time >= p.sim_time) p.trans_time = p.sim_time - p.output;
    if (p.trans_time < 0) p.trans_time = 0;
    if (p.pred_time >= p.sim_time || p.pred_time < 0) p.Npred = 0;
    if (p.Npred == 0) p.pred_time = p.sim_time + 2;
    p.burst_steps = (int)(p.burst_duration / p.dt);
    if (p.BC < 0) p.sizeL = (double)p.N * p.N;
    p.total_outstep = (int)((p.sim_steps - 1) / p.step_output) - (int)(p.trans_time / p.dt) / p.step_output +
                      (0 == std::fmod(p.trans_time / p.dt / p.step_output, 1));
    if (p.total_outstep < 0) p.total_outstep = 0;
    p.total_outstep_pred = p.total_outstep;
    if (p.pred_time > p.sim_time)
        p.total_outstep_pred = 0;
    else if (p.pred_time > p.trans_time)
        p.total_outstep_pred = (int)((p.sim_steps - 1) / p.step_output) -
                               (int)((int)(p.pred_time / p.dt) / p.step_output) +
                               (int)(0 == std::fmod(p.pred_time / p.dt / p.step_output, 1));
    p.out_extend = (p.fileID != "xx");
    p.out_mean = (p.output_mode == 0 || p.output_mode == 1);
    p.out_particle = (p.output_mode == 1);
    if (p.total_outstep == 0) p.out_mean = p.out_particle = false;
}

void output_parameters(const HostParams &p) {  // input_output.cpp:161-201, same keys and formats
    FILE *fp = fopen((p.location + "parameters.dat").c_str(), "w");
    if (!fp) {
        fprintf(stderr, "swarmdyn: cannot write %sparameters.dat\n", p.location.c_str());
        return;
    }
    fprintf(fp, "[Parameters]\n");
    fprintf(fp, "size:               \t%g\n", p.sizeL);
    fprintf(fp, "particles:          \t%d\n", p.N);
    fprintf(fp, "dt:                 \t%g\n", p.dt);
    fprintf(fp, "time:               \t%g\n", p.sim_time);
    fprintf(fp, "trans_time:         \t%g\n", p.trans_time);
    fprintf(fp, "Dp:                 \t%g\n", p.Dphi);
    fprintf(fp, "rep_range:          \t%g\n", p.rep_range);
    fprintf(fp, "alg_range:          \t%g\n", p.alg_range);
    fprintf(fp, "att_range:          \t%g\n", p.att_range);
    fprintf(fp, "soc_strength:       \t%g\n", p.soc_strength);
    fprintf(fp, "prob_social:       \t%g\n", p.prob_social);
    fprintf(fp, "burst_rate:         \t%g\n", p.burst_rate);
    fprintf(fp, "env_strength:       \t%g\n", p.env_strength);
    fprintf(fp, "beta:               \t%g\n", p.beta);
    fprintf(fp, "Npred:              \t%d\n", p.Npred);
    fprintf(fp, "flee_range:         \t%g\n", p.flee_range);
    fprintf(fp, "pred_time:          \t%g\n", p.pred_time);
    fprintf(fp, "pred_speed0:        \t%g\n", p.pred_speed0);
    fprintf(fp, "output_mode:        \t%d\n", p.output_mode);
    fprintf(fp, "N_confu:            \t%d\n", p.N_confu);
    fprintf(fp, "output:             \t%g\n", p.output);
    fprintf(fp, "BC:                 \t%d\n", p.BC);
    fprintf(fp, "IC:                 \t%d\n", p.IC);
    fprintf(fp, "predator_circle_rad:\t%g\n", p.kill_range);
    fprintf(fp, "pred_kill:          \t%d\n", p.pred_kill);
    fprintf(fp, "pred_move:          \t%d\n", p.pred_move);
    fprintf(fp, "out_h5:             \t%d\n", p.out_h5);
    fprintf(fp, "burst_duration:     \t%g\n", p.burst_duration);
    fprintf(fp, "alphaTurn:          \t%g\n", p.alphaTurn);
    fprintf(fp, "kill_rate:          \t%g\n", p.kill_rate);
    fclose(fp);
}

// ---- initial conditions: ResetSystem's cases (settings.cpp:196-387) are generated on the device by sbc_init_state;
// only IC 99 reads "x y vx vy" per line from <location>initCoord.dat here (cf. settings.cpp:222-236) ----------------
bool load_init_coord(const HostParams &p, std::vector<double> &x, std::vector<double> &y, std::vector<double> &phi,
                     std::vector<double> &vproj) {
    const int N = p.N;
    x.assign(N, 0); y.assign(N, 0); phi.assign(N, 0); vproj.assign(N, 1.0);
    std::ifstream in((p.location + "initCoord.dat").c_str());
    if (!in) { fprintf(stderr, "swarmdyn: IC 99 needs %sinitCoord.dat\n", p.location.c_str()); return false; }
    for (int i = 0; i < N; i++) {
        double vx = 1, vy = 0;
        if (!(in >> x[i] >> y[i] >> vx >> vy)) { fprintf(stderr, "swarmdyn: initCoord.dat has fewer than %d rows\n", N); return false; }
        phi[i] = std::atan2(vy, vx);
        vproj[i] = std::sqrt(vx * vx + vy * vy);
    }
    return true;
}

// ---- output (the observables themselves come from the library: sbc_out_swarm / sbc_out_swarm_fishnet) ----
void write_rows(const std::string &file, const std::vector<std::vector<double>> &data) {  // WriteVector2d :322-339
    if (data.empty()) return;
    std::ofstream out(file.c_str(), std::ios_base::trunc);
    for (const auto &row : data) {
        for (double v : row) out << v << " ";
        out << std::endl;
    }
    out << std::endl;
}

// WriteParticles text branch (swarmdyn.cpp:606-624): rows by id up to the last alive agent, zero rows
// for dead agents in between
void append_particles(const std::string &file, const double *rows, const uint8_t *alive, int n, int width) {
    int last = -1;
    for (int i = 0; i < n; i++) if (!alive || alive[i]) last = i;
    if (last < 0) return;
    std::ofstream out(file.c_str(), std::ios::app);
    for (int i = 0; i <= last; i++) {
        for (int k = 0; k < width; k++) out << ((!alive || alive[i]) ? rows[width * i + k] : 0.0) << " ";
        out << std::endl;
    }
}

#define SBC_CHECK(call)                                                                          \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != 0) {                                                                          \
            fprintf(stderr, "swarmdyn: %s failed (%d): %s\n", #call, rc_, sbc_last_error(h));   \
            return 2;                                                                            \
        }                                                                                        \
    } while (0)

}  // namespace

int main(int argc, char **argv) {
    HostParams p;
    parse_parameters(argc, argv, p);
    init_system_parameters(p);
    output_parameters(p);  // swarmdyn.cpp:41-44
    if (p.seed == 0)
        p.seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
    if (p.N < 1) { fprintf(stderr, "swarmdyn: need -N >= 1\n"); return 1; }
    const int R = p.replicates, N = p.N;

    sbc_params sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.sizeL = p.sizeL; sp.dt = p.dt; sp.beta = p.beta; sp.alphaTurn = p.alphaTurn;
    sp.soc_strength = p.soc_strength; sp.env_strength = p.env_strength; sp.prob_social = p.prob_social;
    sp.burst_rate = p.burst_rate; sp.rep_range = p.rep_range; sp.alg_range = p.alg_range; sp.att_range = p.att_range;
    sp.flee_range = p.flee_range; sp.pred_time = p.pred_time; sp.pred_speed0 = p.pred_speed0;
    sp.kill_range = p.kill_range; sp.kill_rate = p.kill_rate; sp.noisep = p.noisep; sp.sim_time = p.sim_time;
    sp.trans_time = p.trans_time; sp.burst_steps = (uint32_t)p.burst_steps; sp.BC = p.BC; sp.N_confu = p.N_confu;
    sp.pred_kill = p.pred_kill; sp.pred_move = p.pred_move;
    sp.neighbour_mode = p.voronoi ? SBC_NEIGHBOUR_VORONOI : SBC_NEIGHBOUR_METRIC; sp.path = SBC_PATH_AUTO;

    sbc_handle *h = nullptr;
    int rc = sbc_create(&sp, N, p.Npred, R, p.seed, 0, 0, &h);
    if (rc != 0) {
        fprintf(stderr, "swarmdyn: sbc_create failed (%d): %s\n", rc, sbc_last_error(nullptr));
        return 2;
    }
    // initial conditions (InitSystem + ResetSystem, swarmdyn.cpp:47-51): on the device for every replicate at once
    if (p.IC != 99) {
        SBC_CHECK(sbc_init_state(h, p.IC));
    } else {
        std::vector<double> x, y, phi, vp, X, Y, PHI, VP;
        if (!load_init_coord(p, x, y, phi, vp)) { sbc_destroy(h); return 1; }
        for (int r = 0; r < R; r++) {   // every replicate starts from the file's state
            X.insert(X.end(), x.begin(), x.end()); Y.insert(Y.end(), y.begin(), y.end());
            PHI.insert(PHI.end(), phi.begin(), phi.end()); VP.insert(VP.end(), vp.begin(), vp.end());
        }
        SBC_CHECK(sbc_set_state(h, X.data(), Y.data(), nullptr, nullptr, PHI.data(), VP.data(), nullptr, nullptr,
                                nullptr, nullptr, nullptr));
    }
    std::cout << "Go";  // swarmdyn.cpp:61

    auto suffix = [&](int r) { return r == 0 ? p.fileID : p.fileID + "_" + std::to_string(r); };
    std::vector<double> part((size_t)R * N * 7), preds((size_t)R * std::max(p.Npred, 1) * 6);
    std::vector<uint8_t> alive((size_t)R * N);
    std::vector<double> obs16((size_t)R * 16), obs4((size_t)R * 4);
    std::vector<int32_t> n_alive(R), n_dead(R);
    std::vector<std::vector<std::vector<double>>> swarm_rows(R), net_rows(R);
    // -J 1: /part and /pred are collected in memory and written once at the end
    std::vector<std::vector<double>> h5_part(R), h5_pred(R);
    std::vector<int> h5_pred_first(R, -1);
    const int s_trans = (int)(p.trans_time / p.dt), s_pred = (int)(p.pred_time / p.dt);
    int outstep = 0;
    const bool want_output = p.out_mean || p.out_particle;
    int s = 0;
    while (s < p.sim_steps) {
        // next step after which the reference calls Output (swarmdyn.cpp:78): s % step_output == 0, s >= s_trans
        int s_out = p.sim_steps - 1;
        bool is_out = false;
        if (want_output) {
            int cand = std::max(s, s_trans);
            cand = ((cand + p.step_output - 1) / p.step_output) * p.step_output;
            if (cand < p.sim_steps) { s_out = cand; is_out = true; }
        }
        SBC_CHECK(sbc_step(h, s, (int64_t)s_out - s + 1));
        s = s_out + 1;
        SBC_CHECK(sbc_get_counts(h, n_alive.data(), n_dead.data()));
        const bool all_dead = std::all_of(n_alive.begin(), n_alive.end(), [](int v) { return v == 0; });
        if (is_out || all_dead) {
            const bool pred_phase = (s_out >= p.pred_time / p.dt);
            if (p.out_particle || all_dead) SBC_CHECK(sbc_get_particles(h, part.data(), alive.data()));
            if (p.Npred > 0 && p.out_particle) SBC_CHECK(sbc_get_predators(h, preds.data()));
            if (p.out_mean) {   // Out_swarm / Out_swarm_fishNet (swarmdyn.cpp:237-344, :358-410) on the device, every replicate at once
                SBC_CHECK(sbc_out_swarm(h, 0, obs16.data(), nullptr));
                if (pred_phase && p.Npred > 0) SBC_CHECK(sbc_out_swarm_fishnet(h, obs4.data()));
            }
            for (int r = 0; r < R; r++) {
                const double *rows = part.data() + (size_t)r * N * 7;
                const uint8_t *al = alive.data() + (size_t)r * N;
                if (p.out_mean) {
                    swarm_rows[r].push_back(std::vector<double>(obs16.begin() + (size_t)r * 16, obs16.begin() + (size_t)(r + 1) * 16));
                    if (pred_phase) {
                        if (p.Npred > 0) net_rows[r].push_back(std::vector<double>(obs4.begin() + (size_t)r * 4, obs4.begin() + (size_t)(r + 1) * 4));
                        else net_rows[r].push_back({0.0, std::numeric_limits<double>::quiet_NaN(), 0.0, (double)n_dead[r]});
                    }
                }
                if (p.out_particle && !p.out_h5) {
                    append_particles(p.location + "part_" + suffix(r) + ".dat", rows, al, N, 7);
                    if (pred_phase && p.Npred > 0)
                        append_particles(p.location + "pred_" + suffix(r) + ".dat", preds.data() + (size_t)r * p.Npred * 6, nullptr, p.Npred, 6);
                }
                if (p.out_particle && p.out_h5 && outstep < p.total_outstep) {
                    // /part: (total_outstep, N, 7), rows by agent id, dead agents stay zero (swarmdyn.cpp:561-600)
                    if (h5_part[r].empty()) h5_part[r].assign((size_t)p.total_outstep * N * 7, 0.0);
                    std::copy(rows, rows + (size_t)N * 7, h5_part[r].begin() + (size_t)outstep * N * 7);
                    if (pred_phase && p.Npred > 0) {  // /pred: (outsteps left at its first output, Npred, 6)
                        if (h5_pred_first[r] < 0) {
                            h5_pred_first[r] = outstep;
                            h5_pred[r].assign((size_t)(p.total_outstep - outstep) * p.Npred * 6, 0.0);
                        }
                        std::copy(preds.data() + (size_t)r * p.Npred * 6, preds.data() + (size_t)(r + 1) * p.Npred * 6,
                                  h5_pred[r].begin() + (size_t)(outstep - h5_pred_first[r]) * p.Npred * 6);
                    }
                }
            }
            outstep++;
        }
        if (all_dead) break;  // swarmdyn.cpp:72-75
    }
    (void)s_pred;
    for (int r = 0; r < R; r++) {  // DataCreateSaveWrite :453-461 writes once, at the last output step
        if (!p.out_h5) {
            write_rows(p.location + "swarm_" + suffix(r) + ".dat", swarm_rows[r]);
            write_rows(p.location + "swarm_fishNet_" + suffix(r) + ".dat", net_rows[r]);
            continue;
        }
        // out_<fileID>.h5 (h5tools.cpp:181-217): datasets at the root for fileID "xx", else under /<fileID>/
        h5mini::Writer w(p.fileID == "xx" ? "" : p.fileID);
        auto flat = [](const std::vector<std::vector<double>> &rows, size_t n_rows, size_t width) {
            std::vector<double> out(n_rows * width, 0.0);
            for (size_t k = 0; k < rows.size() && k < n_rows; k++) std::copy(rows[k].begin(), rows[k].end(), out.begin() + k * width);
            return out;
        };
        int n_sets = 0;
        if (!h5_part[r].empty()) { w.add("part", {(uint64_t)p.total_outstep, (uint64_t)N, 7}, h5_part[r]); n_sets++; }
        if (!h5_pred[r].empty()) { w.add("pred", {(uint64_t)(p.total_outstep - h5_pred_first[r]), (uint64_t)p.Npred, 6}, h5_pred[r]); n_sets++; }
        if (!swarm_rows[r].empty()) { w.add("swarm", {(uint64_t)p.total_outstep, 16}, flat(swarm_rows[r], p.total_outstep, 16)); n_sets++; }
        if (!net_rows[r].empty()) {
            const size_t n_net = (size_t)p.total_outstep - (swarm_rows[r].size() - net_rows[r].size());
            w.add("swarm_fishNet", {(uint64_t)n_net, 4}, flat(net_rows[r], n_net, 4));
            n_sets++;
        }
        if (n_sets && !w.write(p.location + "out_" + suffix(r) + ".h5"))
            fprintf(stderr, "swarmdyn: cannot write %sout_%s.h5\n", p.location.c_str(), suffix(r).c_str());
    }
    if (outstep == 1) {  // swarmdyn.cpp:93-95 "equilibration run": final positions and velocities
        SBC_CHECK(sbc_get_particles(h, part.data(), alive.data()));
        for (int r = 0; r < R; r++) {
            std::ofstream out((p.location + "final_posvel_" + suffix(r) + ".dat").c_str());
            for (int i = 0; i < N; i++) {
                const double *row = part.data() + ((size_t)r * N + i) * 7;
                out << row[0] << " " << row[1] << " " << row[2] << " " << row[3] << std::endl;
            }
            out << std::endl;
        }
    }
    uint64_t st[8];
    if (sbc_get_stats(h, st) == 0 && getenv("SBC_VERBOSE"))
        fprintf(stderr, "\nswarmdyn: %llu agent-steps, %llu directed pairs, %llu kernel launches\n",
                (unsigned long long)st[5], (unsigned long long)st[3],
                (unsigned long long)(st[0] + st[1] + st[2] + st[7]));
    sbc_destroy(h);
    return 0;
}
