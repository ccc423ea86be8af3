// Oracle shim (test infrastructure only): stand-in for GSL's RNG header so that the
// reference's own translation units compile in place without libgsl.
//
// Two modes per generator object:
//   * stream mode  - std::mt19937, bit-identical to GSL's default mt19937 for a 32-bit seed
//                    (gsl_rng_uniform = gen()/2^32, seed 0 -> 4357 as in GSL);
//   * replay mode  - uniforms are popped from an injected queue (decision-level injection,
//                    SURVEY.md F4/H5); running dry aborts loudly.
#ifndef SBC_ORACLE_SHIM_GSL_RNG_H
#define SBC_ORACLE_SHIM_GSL_RNG_H
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <random>

struct gsl_rng_type { const char *name; };
struct gsl_rng {
    std::mt19937 gen;
    bool replay = false;
    std::deque<double> queue;
    std::deque<double> gauss_queue;  // replay mode: values returned by gsl_ran_gaussian
    unsigned long draws = 0;
};

static const gsl_rng_type sbc_shim_mt19937 = {"mt19937"};
static const gsl_rng_type *gsl_rng_default = &sbc_shim_mt19937;

inline const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }
inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { return new gsl_rng(); }
inline void gsl_rng_free(gsl_rng *r) { delete r; }
inline void gsl_rng_set(gsl_rng *r, unsigned long seed) {
    unsigned long s = seed & 0xffffffffUL;
    if (s == 0) s = 4357;
    r->gen.seed(static_cast<std::mt19937::result_type>(s));
}
inline double gsl_rng_uniform(gsl_rng *r) {
    r->draws++;
    if (r->replay) {
        if (r->queue.empty()) {
            std::fprintf(stderr, "oracle gsl shim: replay queue ran dry\n");
            std::abort();
        }
        double v = r->queue.front();
        r->queue.pop_front();
        return v;
    }
    return r->gen() / 4294967296.0;
}
inline double gsl_rng_uniform_pos(gsl_rng *r) {
    double x;
    do { x = gsl_rng_uniform(r); } while (x == 0);
    return x;
}
#endif
