// sbc_api.cu - implementation of the C ABI declared in include/sbc.h.
// Host side only: owns device memory and the stream, converts the reference's double layout to
// the FP32 SoA state, sequences the kernels of one Step() (swarmdyn.cpp:143-204). No arithmetic
// of the hot path runs on the host; without a usable CUDA device every call fails loudly.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sbc.h"
#include "sbc_kernels.cuh"

struct sbc_handle {
    sbc_params hp;
    DevParams P;
    DevState S;
    int device;
    uint64_t seed, replicate_offset;
    cudaStream_t stream;
    bool own_stream;
    std::string err;
    std::vector<uint32_t> geo_host;
    bool state_set;
    ObsScratch *obs;
    uint8_t *flags_dev;
    double *inj_social_dev;
    float *inj_dir_dev;
    uint32_t *inj_k_dev;
    unsigned long long host_stats[8];
    bool force_multikernel;
    bool force_block;        // test hook: keep swarms of 257..2048 agents on the one-block kernel
    unsigned char *stage_raw;   // device staging of the ABI-layout arrays (state_io.cu)
    StateStage stage;
    DenseScratch dense;
    bool dense_ready;
    CellScratch cells;
    bool cells_ready, use_cells;
    bool cells_valid;        // the cell order on the device belongs to the current agents
    int cells_age, cells_max_age;  // steps since the build / steps one build may serve (Verlet skin)
    int cells_local_max_age;       // what this handle's own agents allow (domain mode: the ranks agree on the minimum)
    int dom_period;                // domain mode: agreed number of refresh exchanges between two full ones
    bool dom_last_full;            // kind of the exchange in flight (pack -> unpack)
    // multi-kernel path bookkeeping
    bool list_valid;               // S.active_list[list_step % 3] holds the rows (bin_step == burst_steps) of step list_step
    bool list_split;               // ... and S.edge_list[list_step % 3] those next to a slab edge (decomposed swarm)
    int64_t list_step;
    bool nxt_synced;               // S.pv_nxt carries the dead / free marks of S.pv (needed before a step writes into it)
    bool cs_valid;                 // S.cs == (cos, sin)(phi): true after set_state / ensemble launches, false after multi-kernel steps
    int pv_parity;                 // which of the two position buffers S.pv currently is (captured graphs bake the pointers)
    cudaStream_t stream2;          // second branch of a captured step (rows kernel next to the bulk update)
    cudaStream_t stream3;          // third branch: a decomposed swarm's refresh exchange of the previous step
    cudaEvent_t ev_fork, ev_join, ev_join2;
    int dom_shift;                 // ctl[18] as last uploaded: exchange number minus step index
    double v0_max;           // largest speed handed in through sbc_set_state
    DomainScratch dom;
    // exchange through peer memory: own inbox [flags 256 B | from-left p0 | from-left p1 | from-right p0 | from-right p1]
    unsigned char *inbox;
    size_t inbox_buf_bytes;
    unsigned char *peer_l, *peer_r;        // the neighbours' inboxes (mapped through CUDA IPC, or plain pointers in-process)
    std::vector<void *> ipc_opened;
    unsigned char *send_own[2];            // this rank's send buffers (left, right)
    int dom_epoch;
    // predators of a decomposed swarm (predator.cu): every rank's predator area, this rank's closest-prey key and candidates
    PredBox pbox;
    bool pbox_ready;
    unsigned long long *pred_key;
    int *pred_cand;
    // CUDA graphs of one multi-kernel step, keyed by MK_* bits | position-buffer parity
    std::map<unsigned, cudaGraphExec_t> graphs;
    uint32_t *step_dev;
    int64_t step_dev_value;
    bool use_graphs;
    unsigned long long *trace_dev;
    double *part_dev;           // [R][N][7] rows of particle::out()
    uint8_t *part_alive_dev;
};

static std::string g_create_error;

#define SBC_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            return SBC_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

static int fail(sbc_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

// Captured graphs bake device pointers (S.gid, dom.ctl, halo_list, inbox addresses of the peers ...) by value:
// whoever changes one of them throws the graphs away; they are re-captured on the next step.
static void invalidate_graphs(sbc_handle *h) {
    if (h->graphs.empty()) return;
    cudaStreamSynchronize(h->stream);   // a replay may still be in flight
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
    h->step_dev_value = -1;
}

static float next_down(float v) { return std::nextafterf(v, -INFINITY); }
static float next_up(float v) { return std::nextafterf(v, INFINITY); }

static void fill_dev_params(const sbc_params &p, uint64_t seed, uint64_t rep_off, DevParams &P) {
    std::memset(&P, 0, sizeof(P));
    P.L = (float)p.sizeL;
    P.invL = (float)(1.0 / p.sizeL);
    P.L_lo = (float)(p.sizeL - (double)(float)p.sizeL);
    P.dt = (float)p.dt;
    P.beta = (float)p.beta;
    P.alphaTurn = (float)p.alphaTurn;
    P.soc = (float)p.soc_strength;
    P.env = (float)p.env_strength;
    P.inv_soc = (float)(1.0 / p.soc_strength);
    P.inv_env = (float)(1.0 / p.env_strength);
    P.inv_log2q = (float)(1.0 / std::log2(1.0 - p.burst_rate * p.dt));
    P.rep = (float)p.rep_range;
    P.alg = (float)p.alg_range;
    P.att = (float)p.att_range;
    const double thr[3] = {p.rep_range, p.alg_range, p.att_range};
    // FP32 rounding band of the squared distance around each squared threshold: relative part
    // from dx, dy, the products and the sum (<= 8 u, doubled for margin); the periodic fold is
    // done on the far coordinate first (sbc_delta_pbc), which leaves only the rounding of the
    // low word of a non-representable box length as an absolute term
    const double u = 1.0 / 16777216.0;
    for (int k = 0; k < 3; k++) {
        const double T = thr[k] * thr[k];
        double e_dx = 0;
        if (p.BC == 0) e_dx = u * std::fabs(p.sizeL - (double)(float)p.sizeL) + 1e-12 * p.sizeL;
        const double delta = 16 * u * T + 4 * thr[k] * e_dx + e_dx * e_dx;
        P.band_lo[k] = next_down((float)(T - delta));
        P.band_hi[k] = next_up((float)(T + delta));
    }
    {   // dense pass (pair_dense.cu): conservative "far" threshold and compactness limit
        const double e_abs = (p.BC == 0) ? 4 * u * p.sizeL : 0.0;
        const double att_far = p.att_range + e_abs;
        P.far2 = next_up((float)(att_far * att_far * (1 + 32 * u)));
        if (P.far2 < P.band_hi[2]) P.far2 = P.band_hi[2];
        P.compact_lim = (float)(p.sizeL / 2 - att_far - 1e-4 * p.sizeL);
    }
    P.rep2 = (float)(thr[0] * thr[0]);
    P.alg2 = (float)(thr[1] * thr[1]);
    P.att2 = (float)(thr[2] * thr[2]);
    P.flee_range = (float)p.flee_range;
    P.inv_flee_range = (float)(1.0 / p.flee_range);
    P.pred_speed = (float)p.pred_speed0;
    P.kill_range = (float)p.kill_range;
    P.kill_rate = (float)p.kill_rate;
    P.noisep = (float)p.noisep;
    P.wall_margin_r = (float)(p.sizeL - 2);
    P.t_burst = (float)(p.dt * p.burst_steps);
    P.dL = p.sizeL;
    P.d_rep = p.rep_range;
    P.d_alg = p.alg_range;
    P.d_att = p.att_range;
    P.d_prob_social = p.prob_social;
    P.d_flee_range = p.flee_range;
    P.d_kill_range = p.kill_range;
    P.burst_steps = p.burst_steps;
    P.max_steps = (uint32_t)(5 / (p.burst_rate * p.dt));
    P.tm_bits = 1;
    while ((1ull << P.tm_bits) <= (unsigned long long)p.burst_steps) P.tm_bits++;
    P.tm_mask = (1u << P.tm_bits) - 1u;
    P.tm_stb_max = (P.tm_bits >= 32) ? 0u : (uint32_t)((1ull << (32 - P.tm_bits)) - 1ull);
    P.BC = p.BC;
    P.N_confu = p.N_confu;
    P.pred_kill = p.pred_kill;
    P.pred_move = p.pred_move;
    P.has_alg_zone = p.alg_range > p.rep_range;
    P.voronoi = (p.neighbour_mode == SBC_NEIGHBOUR_VORONOI) ? 1 : 0;
    P.seed_lo = (uint32_t)seed;
    P.seed_hi = (uint32_t)(seed >> 32);
    P.replicate_offset = (uint32_t)rep_off;
}

// inter-burst inverse-CDF table, see DESIGN.md section 5 (the oracle builds its own copy from the
// same recipe: q_k by repeated IEEE multiplication, thr = floor(2^32 (1 - q_k)))
static std::vector<uint32_t> geometric_table(double burst_rate, double dt) {
    const double p = burst_rate * dt;
    const unsigned max_steps = (unsigned)(5 / p);
    std::vector<uint32_t> thr(max_steps);
    double q = 1.0;
    for (unsigned k = 1; k <= max_steps; k++) {
        q *= (1.0 - p);
        double t = 4294967296.0 * (1.0 - q);
        if (t < 0) t = 0;
        if (t > 4294967295.0) t = 4294967295.0;
        thr[k - 1] = (uint32_t)t;
    }
    return thr;
}

template <class T>
static cudaError_t dev_alloc(T **ptr, size_t n) {
    return cudaMalloc(reinterpret_cast<void **>(ptr), n * sizeof(T));
}

extern "C" {

int sbc_abi_version(void) { return SBC_ABI_VERSION; }

const char *sbc_last_error(const sbc_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int sbc_create(const sbc_params *p, int n_agents, int n_pred, int n_replicates, uint64_t seed,
               uint64_t replicate_offset, int device, sbc_handle **out) {
    if (!p || !out) return fail(nullptr, SBC_ERR_INVALID, "sbc_create: null argument");
    *out = nullptr;
    if (n_agents < 1 || n_replicates < 1 || n_pred < 0)
        return fail(nullptr, SBC_ERR_INVALID, "sbc_create: need n_agents >= 1, n_replicates >= 1, n_pred >= 0");
    if (p->neighbour_mode != SBC_NEIGHBOUR_METRIC && p->neighbour_mode != SBC_NEIGHBOUR_VORONOI)
        return fail(nullptr, SBC_ERR_INVALID, "sbc_create: unknown neighbour_mode");
    if (p->neighbour_mode == SBC_NEIGHBOUR_VORONOI && n_agents > 1024)
        return fail(nullptr, SBC_ERR_INVALID,
                    "sbc_create: the Voronoi neighbour rule (InteractionVoronoiF2F / F2FP) is implemented for swarms of up to "
                    "1024 agents");
    if (!(p->dt > 0) || !(p->burst_rate > 0) || !(p->burst_rate * p->dt < 1))
        return fail(nullptr, SBC_ERR_INVALID, "sbc_create: need dt > 0 and 0 < burst_rate*dt < 1");
    if (!(p->rep_range <= p->alg_range && p->alg_range <= p->att_range))
        return fail(nullptr, SBC_ERR_INVALID, "sbc_create: need rep_range <= alg_range <= att_range");
    if (p->BC == 0 && !(p->att_range < p->sizeL / 2))
        return fail(nullptr, SBC_ERR_INVALID, "sbc_create: periodic box needs att_range < sizeL/2 (minimum image)");
    if (p->BC < -1 || p->BC > 6) return fail(nullptr, SBC_ERR_INVALID, "sbc_create: BC must be -1..6");
    if ((size_t)n_agents * (size_t)n_replicates > (size_t)1 << 30)
        return fail(nullptr, SBC_ERR_INVALID, "sbc_create: more than 2^30 agents per handle");
    {   // both burst timers of an agent share one 32-bit word (bin_step <= burst_steps, steps_till_burst <= max_steps + 1)
        unsigned bits = 1;
        while ((1ull << bits) <= (unsigned long long)p->burst_steps) bits++;
        const double max_stb = 5.0 / (p->burst_rate * p->dt) + 2.0;
        if (bits >= 31 || max_stb >= (double)(1ull << (32 - bits)))
            return fail(nullptr, SBC_ERR_INVALID, "sbc_create: burst_steps and 5/(burst_rate dt) do not fit one 32-bit timer word "
                                                  "(dt far below 1e-5)");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        return fail(nullptr, SBC_ERR_CUDA,
                    std::string("sbc_create: no usable CUDA device (") + cudaGetErrorString(e) +
                        "); libsbc has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, SBC_ERR_INVALID, "sbc_create: bad device index");
    sbc_handle *h = new sbc_handle();
    h->hp = *p;
    h->device = device;
    h->seed = seed;
    h->replicate_offset = replicate_offset;
    h->own_stream = true;
    h->state_set = false;
    h->obs = nullptr;
    h->flags_dev = nullptr;
    h->inj_social_dev = nullptr;
    h->inj_dir_dev = nullptr;
    h->inj_k_dev = nullptr;
    h->force_multikernel = false;
    h->force_block = false;
    h->stage_raw = nullptr;
    h->dense_ready = false;
    h->cells_ready = false;
    h->use_cells = false;
    h->cells_valid = false;
    h->cells_age = 0;
    h->cells_max_age = 1;
    h->cells_local_max_age = 0;
    h->dom_period = 0;
    h->dom_last_full = true;
    h->v0_max = 1.0;
    h->list_valid = false;
    h->list_step = -1;
    h->nxt_synced = false;
    h->cs_valid = true;
    h->pv_parity = 0;
    h->stream2 = h->stream3 = nullptr;
    h->ev_fork = h->ev_join = h->ev_join2 = nullptr;
    h->dom_shift = 0x7fffffff;
    h->step_dev = nullptr;
    h->step_dev_value = -1;
    h->use_graphs = getenv("SBC_NO_GRAPHS") == nullptr;
    h->part_dev = nullptr;
    h->part_alive_dev = nullptr;
    std::memset(h->host_stats, 0, sizeof(h->host_stats));
    std::memset(&h->S, 0, sizeof(h->S));
    fill_dev_params(*p, seed, replicate_offset, h->P);
    h->geo_host = geometric_table(p->burst_rate, p->dt);
    DevState &S = h->S;
    S.N = n_agents;
    S.R = n_replicates;
    S.Npred = n_pred;
    const size_t total = (size_t)n_agents * n_replicates;
    const size_t npt = (size_t)(n_pred > 0 ? n_pred : 1) * n_replicates;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    {   // the rows kernel of a captured step runs on its own branch, ahead of the bulk update in the block scheduler
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        ok = ok && cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, hi) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && dev_alloc(&S.pv, total) == cudaSuccess && dev_alloc(&S.pv_dead, total) == cudaSuccess;
    ok = ok && dev_alloc(&S.pv_nxt, total) == cudaSuccess && dev_alloc(&S.cs, total) == cudaSuccess;
    ok = ok && dev_alloc(&S.hv, total) == cudaSuccess && dev_alloc(&S.force, total) == cudaSuccess;
    ok = ok && dev_alloc(&S.tm, total + 4) == cudaSuccess && dev_alloc(&S.alive, total + 4) == cudaSuccess;
    ok = ok && dev_alloc(&S.tm_nxt, total + 4) == cudaSuccess;
    if (n_pred > 1) ok = ok && dev_alloc(&S.kill_list, total) == cudaSuccess && dev_alloc(&S.kill_count, 1) == cudaSuccess;
    ok = ok && dev_alloc(&S.acc, total * 8) == cudaSuccess && dev_alloc(&S.cnt, total * 4) == cudaSuccess;
    ok = ok && dev_alloc(&S.pred_pv, npt) == cudaSuccess && dev_alloc(&S.pred_phi, npt) == cudaSuccess;
    ok = ok && dev_alloc(&S.n_dead, (size_t)n_replicates) == cudaSuccess;
    ok = ok && dev_alloc(&S.pred_spawned, (size_t)n_replicates) == cudaSuccess;
    for (int k = 0; k < 3; k++) ok = ok && dev_alloc(&S.active_list[k], total) == cudaSuccess;
    ok = ok && dev_alloc(&S.active_count, 8) == cudaSuccess;
    ok = ok && dev_alloc(&S.stats, 8) == cudaSuccess;
    if (n_agents <= 32 && n_pred == 0) {   // ensemble kernel in queue mode: one counter, one word per group of tanks
        ok = ok && dev_alloc(&S.ens_queue, 1) == cudaSuccess && dev_alloc(&S.ens_done, (size_t)n_replicates) == cudaSuccess;
    }
    uint32_t *geo_dev = nullptr;
    ok = ok && dev_alloc(&geo_dev, h->geo_host.size() ? h->geo_host.size() : 1) == cudaSuccess;
    double *eff_dev = nullptr;
    ok = ok && dev_alloc(&eff_dev, (size_t)n_agents + 1) == cudaSuccess;
    if (!ok) {
        g_create_error = std::string("sbc_create: CUDA allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        sbc_destroy(h);
        return SBC_ERR_NOMEM;
    }
    S.geo = geo_dev;
    {   // kill_rate * SigmoidFunction(N_sensed, N_confu, -1) (mathtools.cpp:22-33) for every possible N_sensed
        std::vector<double> eff((size_t)n_agents + 1, 0.0);
        for (int n = 1; n <= n_agents; n++)
            eff[n] = p->kill_rate * (0.5 * (std::tanh(-1.0 * ((double)n - (double)p->N_confu)) + 1.0));
        cudaMemcpy(eff_dev, eff.data(), eff.size() * sizeof(double), cudaMemcpyHostToDevice);
        S.kill_eff = eff_dev;
    }
    cudaMemcpy(geo_dev, h->geo_host.data(), h->geo_host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    cudaMemset(S.stats, 0, 8 * sizeof(unsigned long long));
    cudaMemset(S.active_count, 0, 8 * sizeof(int));
    if (S.kill_count) cudaMemset(S.kill_count, 0, sizeof(int));
    cudaMemset(S.n_dead, 0, n_replicates * sizeof(int));
    cudaMemset(S.pred_spawned, 0, n_replicates * sizeof(int));
    cudaMemset(S.acc, 0, total * 8 * sizeof(float));
    cudaMemset(S.cnt, 0, total * 4 * sizeof(int));
    cudaMemset(S.pv_dead, 0, total * sizeof(float4));
    {   // InitPredator defaults (settings.cpp:147-166): x = v = 1, phi = 1
        std::vector<float4> pp(npt, make_float4(1.f, 1.f, 1.f, 1.f));
        std::vector<float> ph(npt, 1.f);
        cudaMemcpy(S.pred_pv, pp.data(), npt * sizeof(float4), cudaMemcpyHostToDevice);
        cudaMemcpy(S.pred_phi, ph.data(), npt * sizeof(float), cudaMemcpyHostToDevice);
    }
    // pair-search path for the burst-gated pass (sbc_params.path)
    if (p->path == SBC_PATH_CELLLIST) {
        if (!sbc_cells_supported(h->P, n_agents, n_replicates)) {
            g_create_error = "sbc_create: SBC_PATH_CELLLIST needs one swarm and, when periodic, sizeL >= 3 att_range";
            sbc_destroy(h);
            return SBC_ERR_INVALID;
        }
        h->use_cells = true;
    } else if (p->path == SBC_PATH_AUTO) {
        h->use_cells = n_agents > 2048 && sbc_cells_supported(h->P, n_agents, n_replicates);
    }
    *out = h;
    // InitSystem defaults (settings.cpp:168-193)
    return sbc_set_state(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr);
}

int sbc_destroy(sbc_handle *h) {
    if (!h) return SBC_OK;
    cudaSetDevice(h->device);
    DevState &S = h->S;
    cudaFree(S.pv); cudaFree(S.pv_nxt); cudaFree(S.pv_dead); cudaFree(S.hv); cudaFree(S.cs); cudaFree(S.force); cudaFree(S.tm); cudaFree(S.tm_nxt);
    cudaFree(S.kill_list); cudaFree(S.kill_count);
    cudaFree(S.alive); cudaFree(S.acc); cudaFree(S.cnt); cudaFree(S.pred_pv); cudaFree(S.pred_phi);
    cudaFree(S.n_dead); cudaFree(S.pred_spawned); cudaFree(S.active_list[0]); cudaFree(S.active_list[1]); cudaFree(S.active_list[2]); cudaFree(S.active_count);
    cudaFree(S.edge_list[0]); cudaFree(S.edge_list[1]); cudaFree(S.edge_list[2]);
    cudaFree(S.ens_queue); cudaFree(S.ens_done);
    cudaFree(S.stats); cudaFree(const_cast<uint32_t *>(S.geo)); cudaFree(const_cast<double *>(S.kill_eff));
    sbc_obs_free(h->obs); cudaFree(h->flags_dev);
    if (h->dense_ready) sbc_dense_scratch_free(h->dense);
    if (h->cells_ready) sbc_cell_scratch_free(h->cells);
    if (h->dom.enabled) sbc_domain_scratch_free(h->dom);
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->ev_join2) cudaEventDestroy(h->ev_join2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(h->pred_key); cudaFree(h->pred_cand); cudaFree(h->inbox); cudaFree(h->send_own[0]); cudaFree(h->send_own[1]);
    cudaFree(h->S.gid);
    cudaFree(h->step_dev);
    cudaFree(h->trace_dev);
    cudaFree(h->stage_raw); cudaFree(h->part_dev); cudaFree(h->part_alive_dev);
    cudaFree(h->inj_social_dev); cudaFree(h->inj_dir_dev); cudaFree(h->inj_k_dev);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SBC_OK;
}

int sbc_set_stream(sbc_handle *h, void *cuda_stream) {
    if (!h) return SBC_ERR_INVALID;
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
    return SBC_OK;
}

int sbc_synchronize(sbc_handle *h) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    return SBC_OK;
}

static int ensure_dense(sbc_handle *h) {
    if (h->dense_ready) return SBC_OK;
    cudaError_t e = sbc_dense_scratch_alloc(h->dense, h->S.N, h->S.R);
    if (e != cudaSuccess) {
        sbc_dense_scratch_free(h->dense);
        return fail(h, SBC_ERR_NOMEM, std::string("dense pass scratch: ") + cudaGetErrorString(e));
    }
    h->dense_ready = true;
    return SBC_OK;
}

// How many steps one cell build may serve: nobody moves faster than max(initial speed, largest force /
// friction) (vproj' = vproj (1 - beta dt) + f dt), a wall reflection can add twice the overshoot
// (agents_dynamics.cpp:432-470); every agent must stay within skin/2 of where the build saw it.
// skin of the cell list as a fraction of att_range (SBC_SKIN_FRAC overrides the default for tuning runs)
static float cells_skin_for(const sbc_handle *h) {
    float frac = 0.2f;
    if (const char *e = std::getenv("SBC_SKIN_FRAC")) frac = (float)std::atof(e);
    return sbc_cells_skin(h->P, frac * h->P.att);
}
static void set_cells_max_age(sbc_handle *h) {
    const double vmax = std::fmax(h->v0_max, std::fmax(h->hp.soc_strength, h->hp.env_strength) / h->hp.beta);
    // BC 1 puts an agent that left the box at 0.9999 L / 0.0001 (agents_dynamics.cpp:300-356): a jump of up to 1e-4 L
    const double clamp_jump = (h->hp.BC == 1) ? 1.0001e-4 * h->hp.sizeL : 0.0;
    const double per_step = (h->hp.BC == 0 ? 1.0 : 3.0) * vmax * h->hp.dt * 1.01 + clamp_jump;
    const double k = (h->cells.skin > 0.f && per_step > 0) ? std::floor(0.5 * h->cells.skin / per_step) : 0.0;
    h->cells_local_max_age = (int)std::fmin(std::fmax(k, 0.0), 64.0);
    // a decomposed swarm rebuilds on every rank in the same step: the agreed period, never more than is safe here
    h->cells_max_age = h->dom.enabled ? std::min(h->cells_local_max_age, h->dom_period) : h->cells_local_max_age;
}

static int ensure_cells(sbc_handle *h) {
    if (h->cells_ready) return SBC_OK;
    // skin: a fifth of att_range when the box allows; one build then serves several steps
    h->cells.skin = cells_skin_for(h);
    cudaError_t e = sbc_cell_scratch_alloc(h->cells, h->P, h->S.N);
    if (e != cudaSuccess) {
        sbc_cell_scratch_free(h->cells);
        return fail(h, SBC_ERR_NOMEM, std::string("cell list scratch: ") + cudaGetErrorString(e));
    }
    h->cells_ready = true;
    h->S.cell_key = h->cells.key_in;
    invalidate_graphs(h);   // (nothing captured yet on this path, but the pointer is baked into captured steps)
    set_cells_max_age(h);
    return SBC_OK;
}

// slots that can hold an agent the update moves: a slab's owned range, otherwise everything
static int n_update_of(const sbc_handle *h) {
    return h->dom.enabled ? h->dom.own_cap : (int)((size_t)h->S.N * h->S.R);
}

// one three-zone pass over the current state WITHOUT advancing it (test hook, dense-pass benchmark): dense (all rows)
// or the burst-gated rows, compacted into list 0
static int launch_pair_pass(sbc_handle *h, int all_rows, bool resort, bool clear_inactive = false) {
    StepArgs A = {};
    A.n_update = n_update_of(h);
    A.edge_hi = -1;
    if (all_rows && !h->P.voronoi) {
        if (int rc = ensure_dense(h)) return rc;
        if (resort || !h->dense.sorted_valid) {
            sbc_launch_dense_sort(h->P, h->S, h->dense, h->stream);
            h->host_stats[7] += 3;
        }
        sbc_launch_pair_dense(h->P, h->S, h->dense, h->stream);
        h->host_stats[0]++;
        h->host_stats[7] += 2;
        h->host_stats[3] += (unsigned long long)h->S.R * h->S.N * (unsigned long long)(h->S.N - 1);
        return SBC_OK;
    }
    // rows into list 0 (the kernels pick the list by the parity of the step: 0 here)
    h->list_valid = false;
    SBC_CUDA(cudaMemsetAsync(h->S.active_count, 0, 8 * sizeof(int), h->stream));
    const bool every = all_rows != 0;   // Delaunay-filtered candidates: every alive row through the rows kernel
    sbc_launch_compact_rows(h->P, h->S, every || clear_inactive, every, A.n_update, h->S.active_list[0], h->S.active_count, h->stream);
    h->host_stats[7] += (every || clear_inactive) ? 2 : 1;
    if (h->use_cells && !every) {
        if (int rc = ensure_cells(h)) return rc;
        A.fresh_cells = 1;   // the build below comes after the compaction above
        sbc_launch_cell_build(h->P, h->S, h->cells, h->stream);
        h->host_stats[6]++;
        h->host_stats[7] += 4;
        sbc_launch_pair_cells(h->P, h->S, h->cells, A, h->stream);
        h->host_stats[7] += 1;
    } else {
        sbc_launch_pair_rows(h->P, h->S, A, h->stream);
    }
    h->host_stats[0]++;
    return SBC_OK;
}

static int ensure_stage(sbc_handle *h) {
    if (h->stage_raw) return SBC_OK;
    const size_t total = (size_t)h->S.N * h->S.R;
    const size_t bytes = total * (8 * sizeof(double) + 2 * sizeof(uint32_t) + 1) + 256;
    SBC_CUDA(cudaMalloc(reinterpret_cast<void **>(&h->stage_raw), bytes));
    unsigned char *q = h->stage_raw;
    for (int k = 0; k < 8; k++) { h->stage.d[k] = reinterpret_cast<double *>(q); q += total * sizeof(double); }
    for (int k = 0; k < 2; k++) { h->stage.u[k] = reinterpret_cast<uint32_t *>(q); q += total * sizeof(uint32_t); }
    h->stage.al = q;
    return SBC_OK;
}

// shared by the double (reference layout) and the float entry point: `elem` = bytes per real number in the caller's arrays
// second half of every state change: the staged arrays named in `mask` become the device state
static int finish_set_state(sbc_handle *h, unsigned mask, int have_prev) {
    DevState &S = h->S;
    sbc_launch_pack_state(h->P, S, h->stage, mask, have_prev, h->stream);
    h->dense.sorted_valid = false;
    h->cells_valid = false;
    h->list_valid = false;
    h->nxt_synced = true;    // the pack kernel writes both position buffers
    h->cs_valid = true;
    // a decomposed swarm rebuilds on every rank in the same step: a new state voids the agreed period (every exchange is
    // a full one until sbc_domain_rebuild_period is called again on all ranks)
    if (h->dom.enabled) h->dom_period = 0;
    if (h->cells_ready || h->dom.enabled) set_cells_max_age(h);
    if (h->dom.enabled) sbc_launch_domain_reset(S, h->dom, h->stream);
    h->host_stats[7]++;
    // the caller's buffers may be reused as soon as we return
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaGetLastError());
    h->state_set = true;
    return SBC_OK;
}

static int set_state_impl(sbc_handle *h, const void *const src[8], size_t elem, const uint32_t *bin_step, const uint32_t *stb,
                          const uint8_t *alive) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_stage(h)) return rc;
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    unsigned mask = 0;
    for (int k = 0; k < 8; k++)
        if (src[k]) {
            mask |= 1u << k;
            SBC_CUDA(cudaMemcpyAsync(h->stage.d[k], src[k], total * elem, cudaMemcpyHostToDevice, h->stream));
        }
    if (bin_step) {
        mask |= SBC_ST_BIN;
        SBC_CUDA(cudaMemcpyAsync(h->stage.u[0], bin_step, total * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    }
    if (stb) {
        mask |= SBC_ST_STB;
        SBC_CUDA(cudaMemcpyAsync(h->stage.u[1], stb, total * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    }
    if (alive) {
        mask |= SBC_ST_ALIVE;
        SBC_CUDA(cudaMemcpyAsync(h->stage.al, alive, total, cudaMemcpyHostToDevice, h->stream));
    }
    if (h->state_set && !h->cs_valid) sbc_launch_refresh_cs(S, h->stream);   // a partial upload keeps the other fields
    h->stage.f32 = (elem == sizeof(float)) ? 1 : 0;
    // the largest speed bounds how long a cell order stays valid (Verlet skin): only the cell-list paths ask for it
    if (src[5] && (h->use_cells || h->dom.enabled)) {
        for (size_t g = 0; g < total; g++) {
            const double v = (elem == sizeof(float)) ? (double)static_cast<const float *>(src[5])[g]
                                                     : static_cast<const double *>(src[5])[g];
            if (v > h->v0_max) h->v0_max = v;
        }
    }
    return finish_set_state(h, mask, h->state_set ? 1 : 0);
}

int sbc_set_state(sbc_handle *h, const double *x, const double *y, const double *vx,
                  const double *vy, const double *phi, const double *vproj, const double *fx,
                  const double *fy, const uint32_t *bin_step, const uint32_t *stb,
                  const uint8_t *alive) {
    const void *src[8] = {x, y, vx, vy, phi, vproj, fx, fy};
    return set_state_impl(h, src, sizeof(double), bin_step, stb, alive);
}

int sbc_set_state_f32(sbc_handle *h, const float *x, const float *y, const float *vx, const float *vy, const float *phi,
                      const float *vproj, const float *fx, const float *fy, const uint32_t *bin_step, const uint32_t *stb,
                      const uint8_t *alive) {
    const void *src[8] = {x, y, vx, vy, phi, vproj, fx, fy};
    return set_state_impl(h, src, sizeof(float), bin_step, stb, alive);
}

int sbc_init_state(sbc_handle *h, int ic) {
    if (!h) return SBC_ERR_INVALID;
    if (ic == 99) return fail(h, SBC_ERR_INVALID, "sbc_init_state: IC 99 reads a file - load it on the host and call sbc_set_state");
    if (h->dom.enabled) return fail(h, SBC_ERR_INVALID, "sbc_init_state: not available on a slab of a decomposed swarm");
    SBC_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_stage(h)) return rc;
    h->stage.f32 = 0;
    sbc_launch_init_cond(h->P, h->S, ic, h->hp.rep_range, h->stage, h->stream);
    h->host_stats[7]++;
    if (h->v0_max < 1.0) h->v0_max = 1.0;     // unit speed, settings.cpp:368-378
    // have_prev = 0: everything the kernel did not write is InitSystem's default (no force, timers 0, alive)
    return finish_set_state(h, SBC_ST_X | SBC_ST_Y | SBC_ST_VX | SBC_ST_VY | SBC_ST_PHI | SBC_ST_VPROJ, 0);
}

int sbc_get_state(sbc_handle *h, double *x, double *y, double *vx, double *vy, double *phi,
                  double *vproj, double *fx, double *fy, uint32_t *bin_step, uint32_t *stb,
                  uint8_t *alive) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    if (int rc = ensure_stage(h)) return rc;
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    double *ddst[8] = {x, y, vx, vy, phi, vproj, fx, fy};
    unsigned mask = 0;
    for (int k = 0; k < 8; k++) if (ddst[k]) mask |= 1u << k;
    if (bin_step) mask |= SBC_ST_BIN;
    if (stb) mask |= SBC_ST_STB;
    if (alive) mask |= SBC_ST_ALIVE;
    h->stage.f32 = 0;
    sbc_launch_unpack_state(h->P, S, h->stage, mask, h->stream);
    h->host_stats[7]++;
    for (int k = 0; k < 8; k++)
        if (ddst[k]) SBC_CUDA(cudaMemcpyAsync(ddst[k], h->stage.d[k], total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (bin_step) SBC_CUDA(cudaMemcpyAsync(bin_step, h->stage.u[0], total * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (stb) SBC_CUDA(cudaMemcpyAsync(stb, h->stage.u[1], total * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (alive) SBC_CUDA(cudaMemcpyAsync(alive, h->stage.al, total, cudaMemcpyDeviceToHost, h->stream));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaGetLastError());
    return SBC_OK;
}

int sbc_get_particles(sbc_handle *h, double *out, uint8_t *alive) {
    if (!h || !out) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    if (!h->part_dev) {
        SBC_CUDA(dev_alloc(&h->part_dev, total * 7));
        SBC_CUDA(dev_alloc(&h->part_alive_dev, total));
    }
    sbc_launch_particles(S, h->part_dev, h->part_alive_dev, 0, h->stream);
    h->host_stats[7]++;
    SBC_CUDA(cudaMemcpyAsync(out, h->part_dev, total * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (alive) SBC_CUDA(cudaMemcpyAsync(alive, h->part_alive_dev, total, cudaMemcpyDeviceToHost, h->stream));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaGetLastError());
    return SBC_OK;
}

int sbc_get_particles_f32(sbc_handle *h, float *out, uint8_t *alive) {
    if (!h || !out) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    if (!h->part_dev) {
        SBC_CUDA(dev_alloc(&h->part_dev, total * 7));
        SBC_CUDA(dev_alloc(&h->part_alive_dev, total));
    }
    sbc_launch_particles(S, h->part_dev, h->part_alive_dev, 1, h->stream);
    h->host_stats[7]++;
    SBC_CUDA(cudaMemcpyAsync(out, h->part_dev, total * 7 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (alive) SBC_CUDA(cudaMemcpyAsync(alive, h->part_alive_dev, total, cudaMemcpyDeviceToHost, h->stream));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaGetLastError());
    return SBC_OK;
}

int sbc_get_predators(sbc_handle *h, double *out) {
    if (!h || !out) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t n = (size_t)S.Npred * S.R;
    if (n == 0) return SBC_OK;
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<float4> pp(n);
    SBC_CUDA(cudaMemcpy(pp.data(), S.pred_pv, n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t j = 0; j < n; j++) {
        double *o = out + j * 6;
        o[0] = pp[j].x; o[1] = pp[j].y; o[2] = pp[j].z; o[3] = pp[j].w; o[4] = 0; o[5] = 0;
    }
    return SBC_OK;
}

int sbc_set_predators(sbc_handle *h, const double *x, const double *y, const double *phi,
                      int spawned) {
    if (!h || !x || !y || !phi) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t n = (size_t)S.Npred * S.R;
    if (n == 0) return SBC_OK;
    std::vector<float4> pp(n);
    std::vector<float> ph(n);
    for (size_t j = 0; j < n; j++) {
        ph[j] = (float)phi[j];
        pp[j] = make_float4((float)x[j], (float)y[j], h->P.pred_speed * cosf(ph[j]),
                            h->P.pred_speed * sinf(ph[j]));
    }
    std::vector<int> sp(S.R, spawned ? 1 : 0);
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaMemcpy(S.pred_pv, pp.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(S.pred_phi, ph.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(S.pred_spawned, sp.data(), S.R * sizeof(int), cudaMemcpyHostToDevice));
    return SBC_OK;
}

int sbc_get_counts(sbc_handle *h, int32_t *n_alive, int32_t *n_dead) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<int> nd(S.R);
    SBC_CUDA(cudaMemcpy(nd.data(), S.n_dead, S.R * sizeof(int), cudaMemcpyDeviceToHost));
    for (int r = 0; r < S.R; r++) {
        if (n_alive) n_alive[r] = S.N - nd[r];
        if (n_dead) n_dead[r] = nd[r];
    }
    return SBC_OK;
}

// ---- one step of the multi-kernel path ---------------------------------------------------------------------------
// Step() (swarmdyn.cpp:143-204) as TWO independent kernels - the rows kernel (agents that start a burst: pair sweep,
// detection, their own update) and the bulk update of everybody else - plus, when due, the cell build in front, the
// predator move + kill behind, and in a decomposed swarm the receipt of the neighbours' records in front of the rows
// kernel and the dispatch of this rank's records at the end. Replayed from a CUDA graph whose two branches run side
// by side; the graph bakes pointers, so there is one per combination of the MK_* bits and the position-buffer parity.
enum { MK_PARITY = 1, MK_REBUILD = 2, MK_PRED = 4, MK_XCHG = 8, MK_PUSH = 16, MK_SERIAL = 32 };

static void domain_launch_refresh_exchange(sbc_handle *h, const DevState &S, const uint32_t *step_dev, uint32_t step_off,
                                           cudaStream_t st);
static unsigned char *inbox_buf(unsigned char *base, size_t bytes, int side, int parity);
static void domain_launch_publish(sbc_handle *h, const uint32_t *step_dev, uint32_t step_off, cudaStream_t st);

static int ensure_rows(sbc_handle *h, int64_t s) {
    if (h->list_valid && h->list_step == s) return SBC_OK;
    const int par = (int)((uint32_t)s % 3u);
    SBC_CUDA(cudaMemsetAsync(h->S.active_count, 0, 8 * sizeof(int), h->stream));
    sbc_launch_compact_rows(h->P, h->S, false, false, n_update_of(h), h->S.active_list[par], h->S.active_count + par, h->stream);
    h->host_stats[7] += 1;
    h->list_valid = true;
    h->list_step = s;
    h->list_split = false;   // one list: the rows next to a slab edge are not apart
    return SBC_OK;
}

// the kernels of one step in launch order on the state S (positions in S.pv, new ones into S.pv_nxt);
// aux == nullptr: everything on `main`. Three branches that do not depend on each other:
//   aux2 : a decomposed swarm's refresh exchange of the PREVIOUS step (send this rank's edge positions, receive the
//          neighbours' into the ghost slots) - only the rows next to a slab edge wait for it, inside the rows kernel
//   aux  : cell build when due, rows kernel          main : bulk update
static void record_step(sbc_handle *h, const DevState &S, unsigned key, const StepArgs &A, cudaStream_t main, cudaStream_t aux,
                        cudaStream_t aux2) {
    cudaStream_t rows_st = main, x_st = main;
    if (aux) {
        cudaEventRecord(h->ev_fork, main);
        cudaStreamWaitEvent(aux, h->ev_fork, 0);
        rows_st = aux;
        if (key & MK_XCHG) {
            if (key & MK_SERIAL) {   // no row can wait for the ghosts inside the rows kernel: the exchange goes first
                x_st = aux;
            } else {
                cudaStreamWaitEvent(aux2, h->ev_fork, 0);
                x_st = aux2;
            }
        }
    }
    if (key & MK_XCHG) domain_launch_refresh_exchange(h, S, A.step_dev, A.step, x_st);
    // the cell order is only read by the rows kernel: it is built on that branch, next to the bulk update
    if (key & MK_REBUILD) sbc_launch_cell_build(h->P, S, h->cells, rows_st);
    if (h->use_cells) sbc_launch_pair_cells(h->P, S, h->cells, A, rows_st);
    else sbc_launch_pair_rows(h->P, S, A, rows_st);
    sbc_launch_update_bulk(h->P, S, A, main);
    if (aux) {
        cudaEventRecord(h->ev_join, aux);
        cudaStreamWaitEvent(main, h->ev_join, 0);
        if ((key & MK_XCHG) && !(key & MK_SERIAL)) {
            cudaEventRecord(h->ev_join2, aux2);
            cudaStreamWaitEvent(main, h->ev_join2, 0);
        }
    }
    DevState Sn = S;   // what the host sees after the step: the new positions are the current ones
    std::swap(Sn.pv, Sn.pv_nxt);
    std::swap(Sn.tm, Sn.tm_nxt);
    if (key & MK_PRED) {
        if (h->dom.enabled) {   // replicated predator, all-gathers through the predator areas of the inboxes (predator.cu)
            h->pbox.timeout_ns = h->dom.timeout_ns;
            sbc_launch_domain_pred_move_kill(h->P, Sn, h->pbox, h->dom.own_cap, h->dom.ctl, h->pred_key, h->pred_cand, A, main);
        } else {
            sbc_launch_pred_move_kill(h->P, Sn, A.step, A.step_dev, main);
        }
    }
    // (the edge agents delivered their new positions themselves; the neighbours' flags are raised by the first block of
    // the receive kernel that heads the next step - or ends this call)
}

// `count` consecutive steps s, s + 1, ... with the same ingredients (count > 1: no cell build among them) as ONE graph
// launch: inside a graph the next step's kernels follow at node-to-node latency instead of a launch gap.
static int mk_step(sbc_handle *h, int64_t s, bool pred_phase, bool xchg_first, int count = 1, bool push = false) {
    DevState &S = h->S;
    const bool injected = S.inj_social || S.inj_dir || S.inj_k;
    if (h->use_cells) if (int rc = ensure_cells(h)) return rc;
    // cell order: rebuilt when stale (Verlet skin used up) and after a state change
    const bool rebuild = h->use_cells && (!h->cells_valid || h->cells_age > h->cells_max_age);
    // slab with the peer-memory exchange: rows whose 3 x 3 cells can hold a ghost are listed apart (by the update that
    // re-arms them) and swept by their own blocks of the lanes kernel, which wait inside the kernel for the exchange that
    // runs beside the step. Other rows kernels (predator phase, crowded cells) sweep ONE list and cannot wait: a split
    // list is merged again, and the exchange runs before the rows kernel instead of beside it.
    const bool lanes_edge = h->dom.enabled && h->peer_l && h->use_cells && sbc_pair_cells_lanes(h->P, S, h->cells, pred_phase ? 1 : 0);
    if (h->list_valid && h->list_step == s && h->list_split && !lanes_edge) h->list_valid = false;
    if (int rc = ensure_rows(h, s)) return rc;
    const bool wait_in_rows = lanes_edge && xchg_first && h->list_split;
    if (!h->nxt_synced) {
        SBC_CUDA(cudaMemcpyAsync(S.pv_nxt, S.pv, (size_t)S.N * S.R * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
        h->nxt_synced = true;
    }
    const unsigned key = (h->pv_parity ? MK_PARITY : 0) | (rebuild ? MK_REBUILD : 0) | (pred_phase ? MK_PRED : 0) |
                         (xchg_first ? MK_XCHG : 0) | (push ? MK_PUSH : 0) | ((xchg_first && !wait_in_rows) ? MK_SERIAL : 0) |
                         ((unsigned)count << 8);
    StepArgs A = {};
    A.step = (uint32_t)s;
    A.n_update = n_update_of(h);
    A.pred_active = pred_phase ? 1 : 0;
    A.net_kill = (pred_phase && S.Npred > 1 && h->hp.pred_kill != 0) ? 1 : 0;
    A.finish = 1;
    A.fresh_cells = rebuild ? 1 : 0;
    A.edge_hi = -1;
    if (h->use_cells && h->cells.ncx > 0) {   // periodic box: fixed grid
        A.cell_ncx = h->cells.ncx;
        A.cell_inv_w = (float)((double)h->cells.ncx / (double)h->P.L);   // == CellGeom::inv_w (celllist.cu), bit for bit
    }
    if (lanes_edge) {
        const double inv_w = (double)h->cells.ncx / h->hp.sizeL;   // cell columns per unit length
        A.edge_lo = (int)std::floor((double)h->dom.x_lo * inv_w) + 1;
        A.edge_hi = (int)std::floor((double)h->dom.x_hi * inv_w) - 1;
        A.edge_blocks = 148;
        A.wait_ns = h->dom.timeout_ns;
        if (wait_in_rows) A.ghost_ready = h->dom.ctl + 17;
    }
    if (push) {   // the exchange behind this step is a refresh: the agents in the edge bands deliver their new positions
        A.halo_idx = h->dom.halo_idx;
        A.xshift = h->dom.ctl + 18;
        for (int par = 0; par < 2; par++) {
            A.xdst[0][par] = inbox_buf(h->peer_l, h->inbox_buf_bytes, 1, par);   // my left neighbour's "from the right"
            A.xdst[1][par] = inbox_buf(h->peer_r, h->inbox_buf_bytes, 0, par);
        }
        // listed agents lay within `halo` of an edge at the last full exchange and have moved less than skin / 2 since
        const float reach = h->dom.halo + 0.5f * h->cells.skin + 1e-3f * h->dom.halo;
        A.band_lo = h->dom.x_lo + reach;
        A.band_hi = h->dom.x_hi - reach;
    }
    const bool graphable = h->use_graphs && !injected && !S.flags && h->stream != nullptr && h->stream != cudaStreamLegacy &&
                           h->stream != cudaStreamPerThread;  // the default streams cannot be captured
    if (count > 1 && !graphable) return fail(h, SBC_ERR_STATE, "mk_step: a batch needs graph replay");
    if (graphable) {
        if (!h->step_dev) SBC_CUDA(dev_alloc(&h->step_dev, 1));
        if (h->step_dev_value != s) {   // the step index lives on the device and is advanced by the step itself
            const uint32_t sv = (uint32_t)s;
            SBC_CUDA(cudaMemcpyAsync(h->step_dev, &sv, sizeof(sv), cudaMemcpyHostToDevice, h->stream));
            SBC_CUDA(cudaStreamSynchronize(h->stream));  // sv is a stack variable
        }
        auto it = h->graphs.find(key);
        if (it == h->graphs.end()) {
            cudaGraph_t graph;
            A.step_dev = h->step_dev;
            SBC_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            DevState Sc = S;
            for (int k = 0; k < count; k++) {
                A.step = (uint32_t)k;   // offset inside the launch; the base lives in *step_dev
                A.fresh_cells = (rebuild && k == 0) ? 1 : 0;   // only the first step of a batch may build the cell order
                record_step(h, Sc, k == 0 ? key : (key & ~(unsigned)MK_REBUILD), A, h->stream, h->stream2, h->stream3);
                std::swap(Sc.pv, Sc.pv_nxt);
                std::swap(Sc.tm, Sc.tm_nxt);
            }
            sbc_launch_step_advance(h->step_dev, count, h->stream);
            cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            if (ce != cudaSuccess) return fail(h, SBC_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
            cudaGraphExec_t exec;
            // node priorities (sbc_launch_prio) only count when the graph is instantiated with this flag
            ce = cudaGraphInstantiateWithFlags(&exec, graph, cudaGraphInstantiateFlagUseNodePriority);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(h, SBC_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
            it = h->graphs.emplace(key, exec).first;
        }
        SBC_CUDA(cudaGraphLaunch(it->second, h->stream));
        h->step_dev_value = s + count;
    } else {
        record_step(h, S, key, A, h->stream, nullptr, nullptr);
        h->step_dev_value = -1;
    }
    // the new positions are the current ones from here on
    if (count & 1) {
        std::swap(S.pv, S.pv_nxt);
        std::swap(S.tm, S.tm_nxt);
        h->pv_parity ^= 1;
    }
    h->list_step = s + count;    // the steps' kernels appended the agents they re-armed
    h->list_split = lanes_edge;  // ... next to a slab edge apart, if this step listed them so
    h->cs_valid = false;
    h->dense.sorted_valid = false;
    h->host_stats[0] += count;
    h->host_stats[1] += count;
    h->host_stats[7] += (rebuild ? 4 : 0) + count * ((pred_phase ? 1 : 0) + (xchg_first ? 1 : 0));
    if (h->use_cells) {
        if (rebuild) { h->cells_valid = true; h->cells_age = 0; h->host_stats[6]++; }
        h->cells_age += count;   // the agents moved once more per step since the build
    }
    return SBC_OK;
}

// how many of the next steps can go into one graph launch: a build of the cell order may head a batch, none falls
// inside it. (A launch gap costs about a fifth of a step, and the kernels of consecutive steps follow each other faster
// inside a graph: 2.6 us between two steps at four per launch, 1.3 at eight.) Every batch size is a graph of its own
// (about a millisecond to capture and instantiate), so only a few sizes are used: a whole cycle of the cell order (the
// build and every step it serves; in a decomposed swarm the steps between the build and the next full exchange), else
// a power of two. `limit`: the largest size the caller can use.
static int mk_batch(const sbc_handle *h, int64_t limit) {
    if (!h->use_graphs || h->S.flags || h->S.inj_social || h->S.inj_dir || h->S.inj_k) return 1;
    if (h->stream == nullptr || h->stream == cudaStreamLegacy || h->stream == cudaStreamPerThread) return 1;
    static int cap = 0;
    if (!cap) { cap = 32; if (const char *e = std::getenv("SBC_MAX_BATCH")) cap = std::max(1, atoi(e)); }   // (tuning runs)
    int64_t room = std::min<int64_t>(limit, cap);
    if (h->use_cells) {
        if (!h->cells_ready) return 1;
        const bool rebuild = !h->cells_valid || h->cells_age > h->cells_max_age;   // this step rebuilds: age 0 again
        room = std::min<int64_t>(room, (int64_t)h->cells_max_age - (rebuild ? 0 : h->cells_age) + 1);
        const int cycle = h->cells_max_age + 1;
        if (rebuild && room >= cycle) return cycle;
        if (h->dom.enabled && !rebuild && h->cells_age == 1 && room >= cycle - 2 && cycle > 3) return cycle - 2;
    }
    for (int b = 16; b > 1; b >>= 1)
        if (room >= b) return b;
    return 1;
}

// CreatePredator / CreateFishNet at step s (slot 0: first spawn, 1: the net's periodic re-creation)
static void launch_spawn(sbc_handle *h, int64_t s, int slot) {
    if (h->dom.enabled) {
        h->pbox.timeout_ns = h->dom.timeout_ns;
        sbc_launch_domain_pred_spawn(h->P, h->S, h->pbox, h->dom.own_cap, h->dom.ctl, s, slot, h->stream);
    } else {
        sbc_launch_pred_spawn(h->P, h->S, s, slot, h->stream);
    }
}

// predator schedule in step indices (swarmdyn.cpp:153-170): is step s in the predator phase, does it (re)create one?
struct PredSchedule {
    bool active = false;
    double x = 0;      // pred_time / dt
    int needed = 0;    // steps between two re-creations of a fish net
    bool phase(int64_t s) const { return active && (double)s >= x; }
    bool spawn0(int64_t s) const { return phase(s) && (double)(s - 1) < x; }
    bool spawn1(int64_t s, int npred) const { return npred > 1 && phase(s) && needed > 0 && ((int)s - (int)x) % needed == 0; }
};
static PredSchedule pred_schedule(const sbc_handle *h) {
    PredSchedule q;
    const sbc_params &p = h->hp;
    if (h->S.Npred <= 0) return q;
    q.active = true;
    q.x = p.pred_time / p.dt;
    if (h->S.Npred > 1) {
        if (p.pred_speed0 > 0) q.needed = (int)(p.sizeL / 2 / (p.pred_speed0 * p.dt));
        else q.needed = (int)((p.sim_time - p.trans_time) / 4 / p.dt);
    }
    return q;
}

int sbc_step(sbc_handle *h, int64_t s_first, int64_t n_steps) {
    if (!h || n_steps < 0) return SBC_ERR_INVALID;
    if (!h->state_set) return fail(h, SBC_ERR_STATE, "sbc_step: no state");
    if (h->dom.enabled && h->S.Npred > 0 && !h->pbox_ready)
        return fail(h, SBC_ERR_STATE, "sbc_step: the predators of a decomposed swarm need sbc_domain_attach_world");
    if (n_steps == 0) return SBC_OK;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const sbc_params &p = h->hp;
    const double dt = p.dt;
    const bool injected = S.inj_social || S.inj_dir || S.inj_k;
    if (injected && n_steps != 1) return fail(h, SBC_ERR_STATE, "sbc_step: injected decisions hold for one step");
    const bool cluster_ok = !h->force_block && sbc_cluster_supported(h->P, S.N, S.Npred);
    const bool block_path = !h->force_multikernel && (S.N <= 1024 || cluster_ok) && S.Npred <= 1024 &&
                            p.path != SBC_PATH_CELLLIST && !h->dom.enabled;
    // predator schedule in step indices (swarmdyn.cpp:153-170)
    long long s_pred = -1;
    int s_trunc = 0, needed = 0;
    if (S.Npred > 0) {
        const double x = p.pred_time / dt;
        s_pred = (long long)std::floor(x);
        while ((double)s_pred < x) s_pred++;
        if (s_pred < 0) s_pred = 0;
        s_trunc = (int)x;
        if (S.Npred > 1) {
            if (p.pred_speed0 > 0) needed = (int)(p.sizeL / 2 / (p.pred_speed0 * dt));
            else needed = (int)((p.sim_time - p.trans_time) / 4 / dt);
        }
    }
    const bool cluster_path = block_path && cluster_ok;
    if (block_path) {
        if (!h->cs_valid) {   // steps on the multi-kernel path do not maintain the cached heading vector
            sbc_launch_refresh_cs(S, h->stream);
            h->cs_valid = true;
        }
        // one launch covers up to 2^31 steps; split only to keep the step counter in 32 bits
        int64_t done = 0;
        while (done < n_steps) {
            const int64_t chunk = (n_steps - done > (1ll << 30)) ? (1ll << 30) : (n_steps - done);
            if (cluster_path)  // 257..2048 agents: one swarm on a cluster of up to 8 CTAs (DSMEM)
                SBC_CUDA(sbc_launch_swarm_cluster(h->P, S, s_first + done, chunk, h->stream, s_pred, s_trunc, needed));
            else
                sbc_launch_swarm_block(h->P, S, s_first + done, chunk, h->stream, s_pred, s_trunc, needed, h->force_block);
            done += chunk;
            h->host_stats[2]++;
        }
        h->list_valid = false;    // timers changed behind the row lists' back
        h->nxt_synced = false;    // kills and moves went to S.pv only
        h->dense.sorted_valid = false;
    } else {
        for (int64_t s = s_first; s < s_first + n_steps; s++) {
            const bool pred_phase = S.Npred > 0 && (double)s >= p.pred_time / dt;
            if (S.Npred > 0 && (double)s >= p.pred_time / dt && (double)(s - 1) < p.pred_time / dt) {
                launch_spawn(h, s, 0);  // swarmdyn.cpp:153-158
                h->host_stats[7]++;
            }
            if (S.Npred > 1 && pred_phase) {  // swarmdyn.cpp:160-170
                const int since = (int)s - (int)(p.pred_time / dt);
                if (needed > 0 && since % needed == 0) {
                    launch_spawn(h, s, 1);
                    h->host_stats[7]++;
                }
            }
            // plain steps go up to eight to a graph launch: none of them spawns (the spawn kernels above are launched
            // eagerly); a batch that would hold a spawn or a change of phase is halved until it does not
            int count = mk_batch(h, s_first + n_steps - s);
            while (count > 1 && S.Npred > 0) {
                bool clean = true;
                for (int k = 1; k < count && clean; k++) {
                    const int64_t sk = s + k;
                    const bool ph = (double)sk >= p.pred_time / dt;
                    const bool spawn0 = ph && (double)(sk - 1) < p.pred_time / dt;
                    const bool spawn1 = S.Npred > 1 && ph && needed > 0 && ((int)sk - (int)(p.pred_time / dt)) % needed == 0;
                    if (ph != pred_phase || spawn0 || spawn1) clean = false;
                }
                if (clean) break;
                count = mk_batch(h, count - 1);
            }
            if (int rc = mk_step(h, s, pred_phase, false, count)) return rc;
            s += count - 1;
        }
    }
    h->host_stats[5] += (unsigned long long)n_steps * (unsigned long long)S.N * (unsigned long long)S.R;
    S.inj_social = nullptr;
    S.inj_dir = nullptr;
    S.inj_k = nullptr;
    SBC_CUDA(cudaGetLastError());
    return SBC_OK;
}

static int obs_ready(sbc_handle *h, const char *who) {
    if (!h->state_set) return fail(h, SBC_ERR_STATE, std::string(who) + ": no state");
    if (h->dom.enabled) return fail(h, SBC_ERR_INVALID, std::string(who) + ": not available on a slab of a decomposed swarm");
    return SBC_OK;
}

int sbc_observables(sbc_handle *h, double *out) {
    if (!h || !out) return SBC_ERR_INVALID;
    if (int rc = obs_ready(h, "sbc_observables")) return rc;
    SBC_CUDA(cudaSetDevice(h->device));
    h->host_stats[7] += 8;
    return sbc_obs_out_swarm(h->P, h->S, h->hp, &h->obs, 0, nullptr, nullptr, out, h->stream, h->err);
}

int sbc_out_swarm(sbc_handle *h, int flags, double *out, double *extra) {
    if (!h || !out) return SBC_ERR_INVALID;
    if (int rc = obs_ready(h, "sbc_out_swarm")) return rc;
    SBC_CUDA(cudaSetDevice(h->device));
    h->host_stats[7] += (flags & SBC_OBS_NO_PAIRS) ? 7 : 8;
    return sbc_obs_out_swarm(h->P, h->S, h->hp, &h->obs, flags, out, extra, nullptr, h->stream, h->err);
}

int sbc_out_swarm_fishnet(sbc_handle *h, double *out) {
    if (!h || !out) return SBC_ERR_INVALID;
    if (int rc = obs_ready(h, "sbc_out_swarm_fishnet")) return rc;
    if (h->S.Npred < 1) return fail(h, SBC_ERR_INVALID, "sbc_out_swarm_fishnet: the handle has no predators");
    SBC_CUDA(cudaSetDevice(h->device));
    h->host_stats[7] += 4;
    return sbc_obs_fishnet(h->P, h->S, &h->obs, out, h->stream, h->err);
}

int sbc_pair_interactions(sbc_handle *h, int flags, int32_t *c_rep, int32_t *c_alg, int32_t *c_att,
                          double *f_rep, double *f_alg, double *f_att, int64_t *nn_ptr,
                          int32_t *nn_idx, int64_t nn_capacity) {
    if (!h) return SBC_ERR_INVALID;
    if (!h->state_set) return fail(h, SBC_ERR_STATE, "sbc_pair_interactions: no state");
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    const int all_rows = (flags & SBC_PAIR_ALL_ROWS) != 0;
    if (int rc = launch_pair_pass(h, all_rows, true, true)) return rc;
    if (!all_rows && h->use_cells) { h->cells_valid = true; h->cells_age = 0; }
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaGetLastError());
    std::vector<float> acc(total * 8);
    std::vector<int> cnt(total * 4);
    SBC_CUDA(cudaMemcpy(acc.data(), S.acc, total * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    SBC_CUDA(cudaMemcpy(cnt.data(), S.cnt, total * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t g = 0; g < total; g++) {
        if (c_rep) c_rep[g] = cnt[g * 4 + 0];
        if (c_alg) c_alg[g] = cnt[g * 4 + 1];
        if (c_att) c_att[g] = cnt[g * 4 + 2];
        if (f_rep) { f_rep[2 * g] = acc[g * 8 + 0]; f_rep[2 * g + 1] = acc[g * 8 + 1]; }
        if (f_alg) { f_alg[2 * g] = acc[g * 8 + 2]; f_alg[2 * g + 1] = acc[g * 8 + 3]; }
        if (f_att) { f_att[2 * g] = acc[g * 8 + 4]; f_att[2 * g + 1] = acc[g * 8 + 5]; }
    }
    if (nn_ptr) {
        std::vector<long long> ptr(total + 1);
        long long run = 0;
        for (size_t g = 0; g < total; g++) {
            ptr[g] = run;
            run += cnt[g * 4 + 0] + cnt[g * 4 + 1] + cnt[g * 4 + 2];
        }
        ptr[total] = run;
        for (size_t g = 0; g <= total; g++) nn_ptr[g] = ptr[g];
        if (nn_idx) {
            if (run > nn_capacity) return fail(h, SBC_ERR_NOMEM, "sbc_pair_interactions: nn_capacity too small");
            long long *ptr_dev = nullptr;
            int *idx_dev = nullptr;
            SBC_CUDA(dev_alloc(&ptr_dev, total + 1));
            SBC_CUDA(dev_alloc(&idx_dev, (size_t)(run > 0 ? run : 1)));
            SBC_CUDA(cudaMemcpy(ptr_dev, ptr.data(), (total + 1) * sizeof(long long), cudaMemcpyHostToDevice));
            sbc_launch_nn_fill(h->P, S, all_rows, ptr_dev, idx_dev, h->stream);
            h->host_stats[7]++;
            SBC_CUDA(cudaStreamSynchronize(h->stream));
            if (run > 0) SBC_CUDA(cudaMemcpy(nn_idx, idx_dev, (size_t)run * sizeof(int), cudaMemcpyDeviceToHost));
            cudaFree(ptr_dev);
            cudaFree(idx_dev);
        }
    }
    return SBC_OK;
}

int sbc_inject_decisions(sbc_handle *h, const double *u_social, const double *u_dir,
                         const uint32_t *k_next) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    if (u_social) {
        if (!h->inj_social_dev) SBC_CUDA(dev_alloc(&h->inj_social_dev, total));
        SBC_CUDA(cudaMemcpy(h->inj_social_dev, u_social, total * sizeof(double), cudaMemcpyHostToDevice));
        S.inj_social = h->inj_social_dev;
    }
    if (u_dir) {
        if (!h->inj_dir_dev) SBC_CUDA(dev_alloc(&h->inj_dir_dev, total));
        std::vector<float> f(total);
        for (size_t g = 0; g < total; g++) f[g] = (float)u_dir[g];
        SBC_CUDA(cudaMemcpy(h->inj_dir_dev, f.data(), total * sizeof(float), cudaMemcpyHostToDevice));
        S.inj_dir = h->inj_dir_dev;
    }
    if (k_next) {
        if (!h->inj_k_dev) SBC_CUDA(dev_alloc(&h->inj_k_dev, total));
        SBC_CUDA(cudaMemcpy(h->inj_k_dev, k_next, total * sizeof(uint32_t), cudaMemcpyHostToDevice));
        S.inj_k = h->inj_k_dev;
    }
    return SBC_OK;
}

int sbc_philox_decisions(sbc_handle *h, int64_t step, double *u_social, double *u_dir,
                         uint32_t *k_next) {
    if (!h) return SBC_ERR_INVALID;
    const DevState &S = h->S;
    for (int r = 0; r < S.R; r++)
        for (int i = 0; i < S.N; i++) {
            const uint4 q = sbc_philox4x32_10((uint32_t)i, (uint32_t)step, SBC_SLOT_DECISION,
                                              (uint32_t)r + h->P.replicate_offset, h->P.seed_lo, h->P.seed_hi);
            const size_t g = (size_t)r * S.N + i;
            if (u_social) u_social[g] = (double)q.x * (1.0 / 4294967296.0);
            if (u_dir) u_dir[g] = (double)((float)(q.y >> 8) * (1.0f / 16777216.0f));
            if (k_next) {
                uint32_t lo = 0, hi = (uint32_t)h->geo_host.size();
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (q.z < h->geo_host[mid]) hi = mid; else lo = mid + 1;
                }
                k_next[g] = lo + 1;
            }
        }
    return SBC_OK;
}

int sbc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint4 q = sbc_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
    return SBC_OK;
}

/* test hook: collect branch flags (OR over steps) into a device buffer; flags==NULL reads+clears */
int sbc_debug_flags(sbc_handle *h, int enable, uint8_t *out) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    DevState &S = h->S;
    const size_t total = (size_t)S.N * S.R;
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    if (enable && !h->flags_dev) {
        SBC_CUDA(dev_alloc(&h->flags_dev, total));
        SBC_CUDA(cudaMemset(h->flags_dev, 0, total));
    }
    if (out && h->flags_dev) {
        SBC_CUDA(cudaMemcpy(out, h->flags_dev, total, cudaMemcpyDeviceToHost));
        SBC_CUDA(cudaMemset(h->flags_dev, 0, total));
    }
    S.flags = enable ? h->flags_dev : nullptr;
    return SBC_OK;
}

/* profiling hook: timeline of the multi-kernel step (see include/sbc.h) */
int sbc_debug_trace(sbc_handle *h, int enable, uint64_t *out) {
    if (!h) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)SBC_TRACE_STEPS * SBC_TR_KERNELS * 2;
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    if (out && h->S.trace) SBC_CUDA(cudaMemcpy(out, h->S.trace, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (enable) {
        if (!h->trace_dev) SBC_CUDA(dev_alloc(&h->trace_dev, n));
        std::vector<unsigned long long> init(n);
        for (size_t k = 0; k < n; k++) init[k] = (k & 1) ? 0ull : ~0ull;   // min of the entries, max of the exits
        SBC_CUDA(cudaMemcpy(h->trace_dev, init.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    unsigned long long *want = enable ? h->trace_dev : nullptr;
    if (want != h->S.trace) {
        invalidate_graphs(h);   // the pointer is baked into captured steps
        h->S.trace = want;
    }
    return SBC_OK;
}

/* test hook: route swarms of <= 1024 agents through the multi-kernel path as well */
int sbc_debug_force_multikernel(sbc_handle *h, int on) {
    if (!h) return SBC_ERR_INVALID;
    h->force_multikernel = (on == 1);
    h->force_block = (on == 2);
    return SBC_OK;
}

int sbc_get_stats(sbc_handle *h, uint64_t out[8]) {
    if (!h || !out) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    unsigned long long dev[8];
    SBC_CUDA(cudaMemcpy(dev, h->S.stats, sizeof(dev), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; k++) out[k] = h->host_stats[k] + dev[k];
    return SBC_OK;
}

int sbc_bench_pair_pass(sbc_handle *h, int flags, int iters, float *ms) {
    if (!h || !ms || iters < 1) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    cudaEvent_t e0, e1;
    SBC_CUDA(cudaEventCreate(&e0));
    SBC_CUDA(cudaEventCreate(&e1));
    const int all_rows = (flags & SBC_PAIR_ALL_ROWS) != 0;
    const bool resort = (flags & SBC_PAIR_RESORT) != 0;
    if (int rc = launch_pair_pass(h, all_rows, true)) return rc;  // warm-up, builds the Morton order
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    SBC_CUDA(cudaEventRecord(e0, h->stream));
    for (int k = 0; k < iters; k++)
        if (int rc = launch_pair_pass(h, all_rows, resort)) return rc;
    SBC_CUDA(cudaEventRecord(e1, h->stream));
    SBC_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    SBC_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = t / (float)iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SBC_CUDA(cudaGetLastError());
    return SBC_OK;
}

/* ---- domain decomposition (see include/sbc.h) ------------------------------------------------- */
int sbc_set_global_ids(sbc_handle *h, const uint32_t *gid) {
    if (!h || !gid) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    const size_t total = (size_t)h->S.N * h->S.R;
    if (!h->S.gid) {
        invalidate_graphs(h);   // graphs captured so far keyed the Philox draws by slot (S.gid == nullptr baked in)
        SBC_CUDA(dev_alloc(&h->S.gid, total));
    }
    SBC_CUDA(cudaMemcpyAsync(h->S.gid, gid, total * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    return SBC_OK;
}

int sbc_domain_config(sbc_handle *h, int rank, int world, int own_capacity, int max_migrants, int max_halo) {
    if (!h || world < 1 || rank < 0 || rank >= world) return SBC_ERR_INVALID;
    if (h->S.R != 1) return fail(h, SBC_ERR_INVALID, "sbc_domain_config: one swarm per handle");
    if (h->hp.BC != 0) return fail(h, SBC_ERR_INVALID, "sbc_domain_config: periodic boundary (BC 0) only");
    if (h->S.Npred > 0 && world > 16) return fail(h, SBC_ERR_INVALID, "sbc_domain_config: predators need world <= 16");
    if (own_capacity < 1 || own_capacity > h->S.N || max_migrants < 1 || max_halo < 1)
        return fail(h, SBC_ERR_INVALID, "sbc_domain_config: bad capacities");
    const float skin = cells_skin_for(h);
    const double w = h->hp.sizeL / world, halo = (h->hp.att_range + (double)skin) * 1.0001;
    if (world > 1 && w < (world == 2 ? 2.1 * halo : 1.05 * halo))
        return fail(h, SBC_ERR_INVALID, "sbc_domain_config: slabs narrower than the halo");
    if (!h->use_cells) return fail(h, SBC_ERR_INVALID, "sbc_domain_config: needs the cell-list path");
    SBC_CUDA(cudaSetDevice(h->device));
    if (h->inbox)   // the neighbours hold this handle's exchange count and buffer sizes: they cannot be reset from here
        return fail(h, SBC_ERR_STATE, "sbc_domain_config: the inbox of this handle is already exported; create a new handle");
    invalidate_graphs(h);
    const unsigned long long keep_timeout = h->dom.timeout_ns;
    if (h->dom.enabled) sbc_domain_scratch_free(h->dom);
    cudaError_t e = sbc_domain_scratch_alloc(h->dom, own_capacity, max_migrants, max_halo, h->S.N);
    h->dom.timeout_ns = keep_timeout;
    if (const char *ev = std::getenv("SBC_DOMAIN_TIMEOUT_S")) {
        const double sec = std::atof(ev);
        if (sec > 0) h->dom.timeout_ns = (unsigned long long)(sec * 1e9);
    }
    if (e != cudaSuccess) return fail(h, SBC_ERR_NOMEM, std::string("domain scratch: ") + cudaGetErrorString(e));
    h->dom.enabled = true;
    for (int k = 0; k < 3; k++)   // rows next to a slab edge are listed apart (worst case: every owned agent)
        if (!h->S.edge_list[k]) SBC_CUDA(dev_alloc(&h->S.edge_list[k], (size_t)h->S.N));
    h->dom.halo = (float)halo;
    h->dom_period = 0;
    h->dom_last_full = true;
    h->cells.skin = skin;
    set_cells_max_age(h);
    h->dom.own_cap = own_capacity;
    h->dom.max_mig = max_migrants;
    h->dom.max_halo = max_halo;
    h->pbox.world = world;
    h->pbox.rank = rank;
    if (h->S.Npred > 0 && !h->pred_key) {
        SBC_CUDA(dev_alloc(&h->pred_key, 1));
        SBC_CUDA(cudaMemset(h->pred_key, 0xff, sizeof(unsigned long long)));
        SBC_CUDA(dev_alloc(&h->pred_cand, (size_t)h->S.N));
    }
    h->dom.x_lo = (float)(w * rank);
    h->dom.x_hi = (rank == world - 1) ? (float)h->hp.sizeL : (float)(w * (rank + 1));
    if (!h->S.gid) {  // default ids = slot index
        std::vector<uint32_t> ids((size_t)h->S.N);
        for (size_t k = 0; k < ids.size(); k++) ids[k] = (uint32_t)k;
        if (int rc = sbc_set_global_ids(h, ids.data())) return rc;
    }
    sbc_launch_domain_reset(h->S, h->dom, h->stream);
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    return SBC_OK;
}

size_t sbc_domain_buffer_bytes(int max_migrants, int max_halo) {
    return (size_t)(1 + max_migrants + max_halo) * sbc_domain_record_bytes();
}

int sbc_domain_pack(sbc_handle *h, void *send_left_dev, void *send_right_dev) {
    if (!h || !h->dom.enabled || !send_left_dev || !send_right_dev) return SBC_ERR_STATE;
    SBC_CUDA(cudaSetDevice(h->device));
    // the kind of this exchange follows the cell list: full (migration + new ghost sets) when the next step
    // rebuilds the cell order, otherwise a position refresh of the ghosts already in place
    h->dom_last_full = !h->cells_ready || !h->cells_valid || h->cells_age > h->cells_max_age;
    if (h->dom_last_full) {
        sbc_launch_domain_pack(h->P, h->S, h->dom, send_left_dev, send_right_dev, h->stream);
        h->nxt_synced = false;   // leavers' slots were freed in S.pv only
        h->list_valid = false;
        h->host_stats[7] += 3;
    } else {
        sbc_launch_domain_refresh_pack(h->S, h->dom, send_left_dev, send_right_dev, h->stream);
        h->host_stats[7] += 1;
    }
    return SBC_OK;
}
int sbc_domain_rebuild_period(sbc_handle *h, int period) {
    if (!h || !h->dom.enabled) return SBC_ERR_STATE;
    if (period >= 0) {
        h->dom_period = period;
        set_cells_max_age(h);
        h->cells_valid = false;   // the next exchange is a full one on every rank
        return h->cells_max_age;
    }
    return h->cells_local_max_age;
}
// ---- exchange through peer memory -----------------------------------------------------------------
static unsigned char *inbox_buf(unsigned char *base, size_t bytes, int side, int parity) {
    return base + 256 + sbc_pred_area_bytes() + (size_t)(side * 2 + parity) * bytes;   // [flags | predator area | 4 buffers]
}
static int *inbox_flag(unsigned char *base, int side, int parity) { return reinterpret_cast<int *>(base) + side * 2 + parity; }

int sbc_domain_inbox(sbc_handle *h, void **base_dev, size_t *bytes, void *ipc_handle_64) {
    if (!h || !h->dom.enabled) return SBC_ERR_STATE;
    SBC_CUDA(cudaSetDevice(h->device));
    if (!h->inbox) {
        h->inbox_buf_bytes = sbc_domain_buffer_bytes(h->dom.max_mig, h->dom.max_halo);
        const size_t total = 256 + sbc_pred_area_bytes() + 4 * h->inbox_buf_bytes;
        SBC_CUDA(cudaMalloc((void **)&h->inbox, total));
        SBC_CUDA(cudaMemset(h->inbox, 0, total));
        for (int k = 0; k < 2; k++) {
            SBC_CUDA(cudaMalloc((void **)&h->send_own[k], h->inbox_buf_bytes));
            SBC_CUDA(cudaMemset(h->send_own[k], 0, h->inbox_buf_bytes));
        }
        h->dom_epoch = 0;
    }
    if (base_dev) *base_dev = h->inbox;
    if (bytes) *bytes = 256 + sbc_pred_area_bytes() + 4 * h->inbox_buf_bytes;
    if (ipc_handle_64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        SBC_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(ipc_handle_64), h->inbox));
    }
    return SBC_OK;
}
int sbc_domain_open_peer(sbc_handle *h, const void *ipc_handle_64, void **mapped_dev) {
    if (!h || !ipc_handle_64 || !mapped_dev) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, ipc_handle_64, sizeof(hd));
    void *p = nullptr;
    SBC_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    *mapped_dev = p;
    return SBC_OK;
}
int sbc_domain_attach_peers(sbc_handle *h, void *left_inbox_dev, void *right_inbox_dev) {
    if (!h || !h->dom.enabled || !h->inbox) return SBC_ERR_STATE;
    if (!left_inbox_dev || !right_inbox_dev) return SBC_ERR_INVALID;
    invalidate_graphs(h);
    h->peer_l = static_cast<unsigned char *>(left_inbox_dev);
    h->peer_r = static_cast<unsigned char *>(right_inbox_dev);
    return SBC_OK;
}
int sbc_domain_attach_world(sbc_handle *h, void *const *inboxes_dev, int n) {
    if (!h || !h->dom.enabled || !h->inbox) return SBC_ERR_STATE;
    if (!inboxes_dev || n != h->pbox.world || n > 16) return fail(h, SBC_ERR_INVALID, "sbc_domain_attach_world: one inbox per rank, in rank order");
    invalidate_graphs(h);
    for (int r = 0; r < 16; r++) h->pbox.peer[r] = nullptr;
    for (int r = 0; r < n; r++) {
        if (!inboxes_dev[r]) return fail(h, SBC_ERR_INVALID, "sbc_domain_attach_world: null inbox");
        h->pbox.peer[r] = static_cast<unsigned char *>(inboxes_dev[r]) + 256;
    }
    if (h->pbox.peer[h->pbox.rank] != h->inbox + 256)
        return fail(h, SBC_ERR_INVALID, "sbc_domain_attach_world: entry [rank] must be this handle's own inbox");
    sbc_pred_preload();
    h->pbox_ready = true;
    return SBC_OK;
}

static void domain_launch_push(sbc_handle *h) {
    // what I send left arrives at the left neighbour as data "from its right" (side 1), and vice versa
    void *dst_l[2], *dst_r[2];
    int *flag_l[2], *flag_r[2];
    for (int par = 0; par < 2; par++) {
        dst_l[par] = inbox_buf(h->peer_l, h->inbox_buf_bytes, 1, par);
        dst_r[par] = inbox_buf(h->peer_r, h->inbox_buf_bytes, 0, par);
        flag_l[par] = inbox_flag(h->peer_l, 1, par);
        flag_r[par] = inbox_flag(h->peer_r, 0, par);
    }
    sbc_launch_domain_push(h->dom, h->send_own[0], h->send_own[1], dst_l, dst_r, flag_l, flag_r, h->dom_epoch, h->stream);
}
static void domain_launch_wait(sbc_handle *h) {
    int *flag_a[2] = {inbox_flag(h->inbox, 0, 0), inbox_flag(h->inbox, 0, 1)};
    int *flag_b[2] = {inbox_flag(h->inbox, 1, 0), inbox_flag(h->inbox, 1, 1)};
    sbc_launch_domain_wait(h->dom, flag_a, flag_b, h->dom_epoch, h->stream);
}
static void domain_launch_refresh_unpack(sbc_handle *h) {
    sbc_launch_domain_refresh_unpack(h->S, h->dom, inbox_buf(h->inbox, h->inbox_buf_bytes, 0, 0),
                                     inbox_buf(h->inbox, h->inbox_buf_bytes, 0, 1), inbox_buf(h->inbox, h->inbox_buf_bytes, 1, 0),
                                     inbox_buf(h->inbox, h->inbox_buf_bytes, 1, 1), h->dom_epoch, h->stream);
}
// Turn what the exchange kernels recorded into a status: a wait that timed out (ctl[21]), record overflows and
// header mismatches (ctl[6]) all mean that some rank stepped with wrong ghosts.
static int domain_check(sbc_handle *h, const char *where) {
    int v[24];
    SBC_CUDA(cudaMemcpyAsync(v, h->dom.ctl, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    if (v[21])
        return fail(h, SBC_ERR_STATE, std::string(where) + ": a neighbour's records did not arrive within " +
                                          std::to_string(h->dom.timeout_ns / 1000000ull) + " ms (sbc_domain_set_timeout); "
                                          "the exchange was not applied, the state of this handle is no longer valid");
    if (v[6])
        return fail(h, SBC_ERR_STATE, std::string(where) + ": " + std::to_string(v[6]) +
                                          " exchange events overflowed a record buffer or disagreed with the neighbour's header");
    return SBC_OK;
}
int sbc_domain_set_timeout(sbc_handle *h, double seconds) {
    if (!h || !(seconds > 0)) return SBC_ERR_INVALID;
    invalidate_graphs(h);   // the limit is a kernel argument of the captured wait
    h->dom.timeout_ns = (unsigned long long)(seconds * 1e9);
    return SBC_OK;
}

int sbc_domain_post(sbc_handle *h) {
    if (!h || !h->dom.enabled || !h->peer_l || !h->peer_r) return SBC_ERR_STATE;
    if (int rc = sbc_domain_pack(h, h->send_own[0], h->send_own[1])) return rc;
    h->dom_epoch++;   // exchanges are numbered; the neighbours count along
    domain_launch_push(h);
    h->host_stats[7] += 2;
    return SBC_OK;
}
static int domain_complete_async(sbc_handle *h) {
    if (!h || !h->dom.enabled || !h->inbox) return SBC_ERR_STATE;
    SBC_CUDA(cudaSetDevice(h->device));
    domain_launch_wait(h);
    h->host_stats[7] += 1;
    if (!h->dom_last_full) {
        domain_launch_refresh_unpack(h);
        h->host_stats[7] += 1;
        return SBC_OK;
    }
    const int par = h->dom_epoch & 1;
    return sbc_domain_unpack(h, inbox_buf(h->inbox, h->inbox_buf_bytes, 0, par), inbox_buf(h->inbox, h->inbox_buf_bytes, 1, par),
                             h->send_own[0], h->send_own[1]);
}
int sbc_domain_complete(sbc_handle *h) {
    if (int rc = domain_complete_async(h)) return rc;
    // several handles of one process on one stream (loop-back tests) post first and complete afterwards: the status of
    // this handle is known once its own wait ran, which the stream order guarantees here
    return domain_check(h, "sbc_domain_complete");
}
// the two halves of a refresh exchange (domain.cu): SEND gathers the listed positions straight into the neighbours'
// inboxes and publishes the exchange count; RECV holds until both neighbours' counts arrived and scatters their
// records into the ghost slots
static void domain_peer_pointers(sbc_handle *h, void *dst_l[2], void *dst_r[2], int *flag_l[2], int *flag_r[2],
                                 const void *recv_a[2], const void *recv_b[2], int *flag_a[2], int *flag_b[2]) {
    for (int par = 0; par < 2; par++) {
        dst_l[par] = inbox_buf(h->peer_l, h->inbox_buf_bytes, 1, par);
        dst_r[par] = inbox_buf(h->peer_r, h->inbox_buf_bytes, 0, par);
        flag_l[par] = inbox_flag(h->peer_l, 1, par);
        flag_r[par] = inbox_flag(h->peer_r, 0, par);
        recv_a[par] = inbox_buf(h->inbox, h->inbox_buf_bytes, 0, par);
        recv_b[par] = inbox_buf(h->inbox, h->inbox_buf_bytes, 1, par);
        flag_a[par] = inbox_flag(h->inbox, 0, par);
        flag_b[par] = inbox_flag(h->inbox, 1, par);
    }
}
// the refresh exchange of the step BEFORE step (*step_dev + step_off): send, then receive, on one stream
static void domain_launch_refresh_exchange(sbc_handle *h, const DevState &S, const uint32_t *step_dev, uint32_t step_off,
                                           cudaStream_t st) {
    void *dst_l[2], *dst_r[2];
    const void *recv_a[2], *recv_b[2];
    int *flag_l[2], *flag_r[2], *flag_a[2], *flag_b[2];
    domain_peer_pointers(h, dst_l, dst_r, flag_l, flag_r, recv_a, recv_b, flag_a, flag_b);
    sbc_launch_domain_refresh_receive(S, h->dom, step_dev, step_off, recv_a, recv_b, flag_a, flag_b, flag_l, flag_r, st);
}

static void domain_launch_publish(sbc_handle *h, const uint32_t *step_dev, uint32_t step_off, cudaStream_t st) {
    void *dst_l[2], *dst_r[2];
    const void *recv_a[2], *recv_b[2];
    int *flag_l[2], *flag_r[2], *flag_a[2], *flag_b[2];
    domain_peer_pointers(h, dst_l, dst_r, flag_l, flag_r, recv_a, recv_b, flag_a, flag_b);
    sbc_launch_domain_publish(h->dom, step_dev, step_off, flag_l, flag_r, st);
}

// n_steps x { step ; exchange }. A refresh exchange does not sit between two steps: this rank's edge positions of step
// s leave, and the neighbours' arrive, on a branch of step s + 1's graph that runs BESIDE its bulk update and its rows
// kernel; only the rows next to a slab edge hold (inside the rows kernel, right before they read candidate
// positions) until the ghosts are in place. A rank that is a little ahead or behind costs its neighbours nothing.
int sbc_domain_step(sbc_handle *h, int64_t s_first, int64_t n_steps) {
    if (!h || !h->dom.enabled || !h->peer_l || !h->peer_r) return SBC_ERR_STATE;
    if (!h->state_set) return fail(h, SBC_ERR_STATE, "sbc_domain_step: no state");
    SBC_CUDA(cudaSetDevice(h->device));
    if (n_steps <= 0) return SBC_OK;
    {   // exchange number of the exchange behind step s: s + shift (the next one is dom_epoch + 1, behind s_first)
        const int shift = h->dom_epoch + 1 - (int)s_first;
        if (shift != h->dom_shift) {
            SBC_CUDA(cudaMemcpyAsync(h->dom.ctl + 18, &shift, sizeof(int), cudaMemcpyHostToDevice, h->stream));
            SBC_CUDA(cudaStreamSynchronize(h->stream));   // `shift` is a stack variable
            h->dom_shift = shift;
        }
    }
    if (h->S.Npred > 0 && !h->pbox_ready)
        return fail(h, SBC_ERR_STATE, "sbc_domain_step: the predators of a decomposed swarm need sbc_domain_attach_world");
    const PredSchedule sched = pred_schedule(h);
    const int npred = h->S.Npred;
    bool pending = false;   // the refresh exchange behind the last step has not run yet
    int64_t s = s_first;
    for (; s < s_first + n_steps; s++) {
        if (int rc = ensure_cells(h)) return rc;
        const bool pred_phase = sched.phase(s);
        if (sched.spawn0(s)) { launch_spawn(h, s, 0); h->host_stats[7]++; }
        if (sched.spawn1(s, npred)) { launch_spawn(h, s, 1); h->host_stats[7]++; }
        // the exchange behind this step is a full one when the NEXT step rebuilds the cell order
        const bool rebuild_now = !h->cells_valid || h->cells_age > h->cells_max_age;
        // the refresh steps up to the next full exchange in one graph launch
        int count = pending ? mk_batch(h, s_first + n_steps - s) : 1;
        // (the step that ends in a full exchange goes alone; a step that rebuilds the cell order heads no batch here)
        if (count > 1) count = rebuild_now ? 1 : mk_batch(h, std::min<int64_t>(count, h->cells_max_age - h->cells_age));
        while (count > 1) {   // a batch holds steps of one kind, none of which (re)creates a predator: halved until it does
            bool clean = true;
            for (int k = 0; k < count && clean; k++)
                if (sched.phase(s + k) != pred_phase || sched.spawn0(s + k) || sched.spawn1(s + k, npred)) clean = false;
            if (clean) break;
            count = mk_batch(h, count - 1);
        }
        if (count > 1) {
            if (int rc = mk_step(h, s, pred_phase, true, count, true)) return rc;
            h->dom_last_full = false;
            h->dom_epoch += count;
            s += count - 1;
            continue;   // still pending: the exchange behind the last of these steps
        }
        const int age_after = (rebuild_now ? 0 : h->cells_age) + 1;
        const bool full_after = age_after > h->cells_max_age;
        if (int rc = mk_step(h, s, pred_phase, pending, 1, !full_after)) return rc;
        pending = false;
        if (full_after) {
            if (int rc = sbc_domain_post(h)) return rc;
            if (int rc = domain_complete_async(h)) return rc;
        } else {
            h->dom_last_full = false;
            h->dom_epoch++;
            pending = true;
        }
    }
    if (pending) {   // leave the handle with its ghosts in place: the exchange behind the last step, eagerly
        domain_launch_refresh_exchange(h, h->S, nullptr, (uint32_t)s, h->stream);
        h->host_stats[7] += 2;
    }
    h->host_stats[5] += (unsigned long long)n_steps * (unsigned long long)h->S.N;
    // one host synchronisation per call (not per step): did every exchange of this call arrive and fit?
    return domain_check(h, "sbc_domain_step");
}

int sbc_domain_unpack(sbc_handle *h, const void *recv_a_dev, const void *recv_b_dev, const void *sent_left_dev,
                      const void *sent_right_dev) {
    if (!h || !h->dom.enabled) return SBC_ERR_STATE;
    SBC_CUDA(cudaSetDevice(h->device));
    if (h->dom_last_full) {
        sbc_launch_domain_unpack(h->S, h->dom, recv_a_dev, recv_b_dev, sent_left_dev, sent_right_dev,
                                 h->inbox ? h->dom_epoch : -1, h->stream);
        h->cells_valid = false;   // slots changed hands: the next step rebuilds the cell order
        h->nxt_synced = false;    // ... and both position buffers and the row list start afresh
        h->list_valid = false;
        h->host_stats[7] += 2;
    } else {
        sbc_launch_domain_refresh_unpack(h->S, h->dom, recv_a_dev, recv_a_dev, recv_b_dev, recv_b_dev, h->dom_epoch, h->stream);
        h->host_stats[7] += 1;
    }
    return SBC_OK;
}
int sbc_domain_counts(sbc_handle *h, int32_t *n_owned, int32_t *n_ghost, int32_t *overflow_events) {
    if (!h || !h->dom.enabled) return SBC_ERR_STATE;
    SBC_CUDA(cudaSetDevice(h->device));
    sbc_launch_count_owned(h->S, h->dom, h->stream);
    int v[16];
    SBC_CUDA(cudaMemcpyAsync(v, h->dom.ctl, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    if (n_owned) *n_owned = v[8];
    if (n_ghost) *n_ghost = v[9];
    if (overflow_events) *overflow_events = v[6];
    return SBC_OK;
}
int sbc_get_global_ids(sbc_handle *h, uint32_t *gid) {
    if (!h || !gid) return SBC_ERR_INVALID;
    SBC_CUDA(cudaSetDevice(h->device));
    const size_t total = (size_t)h->S.N * h->S.R;
    SBC_CUDA(cudaStreamSynchronize(h->stream));
    if (h->S.gid) {
        SBC_CUDA(cudaMemcpy(gid, h->S.gid, total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    } else {
        for (size_t k = 0; k < total; k++) gid[k] = (uint32_t)(k % h->S.N);
    }
    return SBC_OK;
}

}  // extern "C"
