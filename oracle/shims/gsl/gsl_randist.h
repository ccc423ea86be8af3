// Oracle shim (test infrastructure only): gsl_ran_gaussian by the polar Box-Muller method with
// rejection, as published for GSL's randist/gauss.c.
#ifndef SBC_ORACLE_SHIM_GSL_RANDIST_H
#define SBC_ORACLE_SHIM_GSL_RANDIST_H
#include <cmath>
#include "gsl_rng.h"
inline double gsl_ran_gaussian(gsl_rng *r, double sigma) {
    if (r->replay) {
        if (r->gauss_queue.empty()) {
            std::fprintf(stderr, "oracle gsl shim: gaussian replay queue ran dry\n");
            std::abort();
        }
        const double g = r->gauss_queue.front();
        r->gauss_queue.pop_front();
        return sigma * g;
    }
    double x, y, r2;
    do {
        x = -1 + 2 * gsl_rng_uniform_pos(r);
        y = -1 + 2 * gsl_rng_uniform_pos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return sigma * y * std::sqrt(-2.0 * std::log(r2) / r2);
}
#endif
