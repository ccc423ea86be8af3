/* orc_philox.h - TEST INFRASTRUCTURE ONLY (shared by both oracle builds).
 *
 * Independent host implementation of the repo's decision protocol:
 * Philox4x32-10 (Salmon et al., SC'11; Random123 constants) keyed by the 64-bit seed, counter
 * (agent_id, step, slot, replicate). The product has its own device implementation in
 * sde_burst_coast_b200/csrc/sbc_philox.cuh; tests check the two against each other and against
 * the Random123 known-answer vectors.
 */
#ifndef ORC_PHILOX_H
#define ORC_PHILOX_H
#include <stdint.h>
#include <vector>

struct OrcPhilox4 { uint32_t v[4]; };

static inline OrcPhilox4 orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1) {
    const uint64_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = M0 * c0, p1 = M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    OrcPhilox4 r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

/* slots of the protocol */
enum { ORC_SLOT_DECISION = 0, ORC_SLOT_DETECT0 = 1 };
static const uint32_t ORC_PREDATOR_AGENT_ID = 0xFFFFFFFFu;

static inline OrcPhilox4 orc_draw(uint64_t seed, uint64_t replicate, uint32_t agent, int64_t step,
                                  uint32_t slot) {
    return orc_philox4x32_10(agent, (uint32_t)step, slot, (uint32_t)replicate,
                             (uint32_t)seed, (uint32_t)(seed >> 32));
}

static inline double orc_u32_to_unit(uint32_t r) { return r * (1.0 / 4294967296.0); }
static inline double orc_u24_to_unit(uint32_t r) { return (r >> 8) * (1.0 / 16777216.0); }

/* Inter-burst interval table (restates the rejection loop agents_dynamics.cpp:250-270 as an
 * inverse CDF on a 32-bit integer draw): thr[k-1] = floor(2^32 (1 - q^k)), q = 1 - burst_rate dt,
 * q^k by repeated IEEE multiplication; k = 1 + #{thr <= r}, capped at max_steps + 1 with
 * max_steps = (unsigned)(5 / (burst_rate dt)). */
static inline std::vector<uint32_t> orc_geometric_table(double burst_rate, double dt) {
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
static inline uint32_t orc_geometric_k(const std::vector<uint32_t> &thr, uint32_t r) {
    /* smallest k>=1 with r < thr[k-1]; none -> size+1 */
    size_t lo = 0, hi = thr.size();
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (r < thr[mid]) hi = mid; else lo = mid + 1;
    }
    return (uint32_t)lo + 1;
}

/* ---- initial-condition protocol (ResetSystem, settings.cpp:196-387) ------------------------------
 * The uniforms the reference pops from its serial stream, as words of Philox blocks: counter (loop index i, step
 * 0xFFFFFFFF, slot 16): v[0] -> x draw, v[1] -> y draw, v[2] -> heading draw; counter (0xFFFFFFFE, ., 16): v[0] -> the
 * swarm-wide heading; counter (0xFFFFFFFE, ., 17): round keys of the ring shuffle (four-round Feistel bijection on the
 * next even power of two, cycle-walked into [0, N)). */
static const uint32_t ORC_IC_STEP = 0xFFFFFFFFu, ORC_IC_SWARM_ID = 0xFFFFFFFEu;
static inline uint32_t orc_ic_mix(uint32_t v) {
    v *= 0x9E3779B1u; v ^= v >> 15; v *= 0x85EBCA77u; v ^= v >> 13;
    return v;
}
static inline uint32_t orc_ic_permute(uint32_t i, uint32_t n, const OrcPhilox4 &key) {
    uint32_t bits = 2;
    while ((1ull << bits) < n) bits += 2;
    const uint32_t half = bits / 2, mask = (1u << half) - 1u;
    uint32_t x = i;
    do {
        uint32_t l = x >> half, r = x & mask;
        for (int q = 0; q < 4; q++) { const uint32_t t = l ^ (orc_ic_mix(r ^ key.v[q]) & mask); l = r; r = t; }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}
/* the uniforms of one replicate in the order ResetSystem consumes them for initial condition `ic` */
static inline std::vector<double> orc_ic_uniform_stream(uint64_t seed, uint64_t replicate, int ic, int N) {
    std::vector<double> u;
    const double theta_u = orc_u32_to_unit(orc_draw(seed, replicate, ORC_IC_SWARM_ID, ORC_IC_STEP, 16).v[0]);
    if (ic == 1 || ic == 2 || ic == 3) u.push_back(theta_u);
    for (int i = 0; i < N; i++) {
        const OrcPhilox4 d = orc_draw(seed, replicate, (uint32_t)i, ORC_IC_STEP, 16);
        const double ux = orc_u32_to_unit(d.v[0]), uy = orc_u32_to_unit(d.v[1]), up = orc_u32_to_unit(d.v[2]);
        if (ic == 1 || ic == 2 || ic == 3 || ic == 4) { u.push_back(ux); u.push_back(uy); }
        else if (ic == 5) { u.push_back(ux); u.push_back(uy); u.push_back(up); }
        else { u.push_back(up); u.push_back(ux); u.push_back(uy); }      /* IC 0 and the "else" branch: heading first */
    }
    return u;
}

#endif
