// sbc_oracle.cpp - TEST INFRASTRUCTURE ONLY ("port" oracle, builds liborc_port.so).
//
// Double-precision CPU restatement of the reference's per-timestep hot path in METRIC neighbour
// mode, written for this repo (no reference code is included or copied). Each function names the
// reference lines whose behaviour it restates; operation ORDER is kept identical wherever
// rounding can be observed, because tests demand bit-identical doubles from this file and from
// the reference's own translation units (oracle/_ref, see ref_wrapper.cpp).
//
// Random decisions follow the repo's Philox protocol (orc_philox.h) or are injected per agent.
#include <chrono>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <vector>
#include <limits>

#include "sbc_oracle.h"
#include "orc_philox.h"

namespace {

struct Agent {
    double x[2], v[2], u[2], phi, vproj, force[2];
    double f_rep[2], f_alg[2], f_att[2], f_flee[2];
    int c_rep, c_alg, c_att, c_flee;
    unsigned bin_step, stb;
    bool dead;
    std::vector<int> nn;
};
struct Pred {
    double x[2], v[2], u[2], phi;
};

}  // namespace

struct orc_world {
    sbc_params p;
    int N, Npred;
    uint64_t seed, replicate;
    std::vector<Agent> a;
    std::vector<Pred> preds;
    int n_dead;
    std::vector<uint32_t> geo;
    // one-shot injection
    bool inj_social, inj_dir, inj_k;
    std::vector<double> u_social, u_dir;
    std::vector<uint32_t> k_next;
};

namespace {

const double PI = M_PI;

inline int sgn(double v) { return (0.0 < v) - (v < 0.0); }

// mathtools.h:75-83 vec_length: sequential sum of squares, then sqrt
inline double len2d(const double *v) {
    double len = 0;
    len += v[0] * v[0];
    len += v[1] * v[1];
    return std::sqrt(len);
}
// mathtools.h:172-179 vec_dot
inline double dot2d(const double *a, const double *b) {
    double d = 0;
    d += a[0] * b[0];
    d += a[1] * b[1];
    return d;
}
// mathtools.h:138-144 vec_set_mag (no guard against a zero vector)
inline void set_mag(const double *in, double mag, double *out) {
    const double cur = len2d(in);
    const double corr = mag / cur;
    out[0] = in[0] * corr;
    out[1] = in[1] * corr;
}
// mathtools.h:54-73 CalcDistVec: r = xj - xi, min image only when BC == 0
inline void dist_vec(const double *xi, const double *xj, int BC, double L, double *r) {
    r[0] = xj[0] - xi[0];
    r[1] = xj[1] - xi[1];
    if (!BC) {
        for (int d = 0; d < 2; d++) {
            const double s = sgn(r[d]);
            r[d] = std::fmod(r[d] + s * L / 2, L) - s * L / 2;
        }
    }
}

// agents_dynamics.cpp:276-471 Boundary<agent>, all seven cases; T has x, v, u, phi
template <class T>
void boundary(T &a, double L, int BC) {
    const double dphi = 0.1;
    double tmpphi, dx, tx;
    switch (BC) {
    case 0:  // :292-298
        a.x[0] = std::fmod(a.x[0] + L, L);
        a.x[1] = std::fmod(a.x[1] + L, L);
        break;
    case 1:  // :300-356 inelastic box
        if (a.x[0] > L) {
            a.x[0] = 0.9999 * L;
            tmpphi = (a.v[1] > 0.) ? (0.5 + dphi) * PI : -(0.5 + dphi) * PI;
            a.v[0] = std::cos(tmpphi);
            a.v[1] = std::sin(tmpphi);
        } else if (a.x[0] < 0) {
            a.x[0] = 0.0001;
            tmpphi = (a.v[1] > 0.) ? (0.5 - dphi) * PI : -(0.5 - dphi) * PI;
            a.v[0] = std::cos(tmpphi);
            a.v[1] = std::sin(tmpphi);
        }
        if (a.x[1] > L) {
            a.x[1] = 0.9999 * L;
            tmpphi = (a.v[0] > 0.) ? -dphi : PI + dphi;
            a.v[0] = std::cos(tmpphi);
            a.v[1] = std::sin(tmpphi);
        } else if (a.x[1] < 0) {
            a.x[1] = 0.0001;
            tmpphi = (a.v[0] > 0.) ? dphi : PI - dphi;
            a.v[0] = std::cos(tmpphi);
            a.v[1] = std::sin(tmpphi);
        }
        break;
    case 2:  // :358-384 elastic box
        for (int d = 0; d < 2; d++) {
            if (a.x[d] > L) {
                dx = 2. * (a.x[d] - L);
                a.x[d] -= dx;
                a.v[d] *= -1.;
            } else if (a.x[d] < 0) {
                dx = 2. * a.x[d];
                a.x[d] -= dx;
                a.v[d] *= -1.;
            }
        }
        break;
    case 3:  // :386-406 x periodic, y elastic
        tx = std::fmod(a.x[0], L);
        if (tx < 0.0f) tx += L;
        a.x[0] = tx;
        if (a.x[1] > L) {
            dx = 2. * (a.x[1] - L);
            a.x[1] -= dx;
            a.v[1] *= -1.;
        } else if (a.x[1] < 0) {
            dx = 2. * a.x[1];
            a.x[1] -= dx;
            a.v[1] *= -1.;
        }
        break;
    case 4:  // :407-430 x periodic, y inelastic
        tx = std::fmod(a.x[0], L);
        if (tx < 0.0f) tx += L;
        a.x[0] = tx;
        if (a.x[1] > L) {
            dx = (a.x[1] - L);
            a.x[1] -= dx + 0.0001;
            a.v[1] = 0.0;
        } else if (a.x[1] < 0) {
            dx = a.x[1];
            a.x[1] -= dx - 0.0001;
            a.v[1] = 0.0;
        }
        break;
    case 5:    // :432-450 elastic circle
    case 6: {  // :452-470 half-elastic circle
        const double dist2cen = len2d(a.x);
        double diff = dist2cen - L;
        if (diff > 0) {
            double n[2] = {a.x[0], a.x[1]};
            const double inv = 1 / dist2cen;
            n[0] *= inv;
            n[1] *= inv;
            const double f = -2 * diff;
            a.x[0] += n[0] * f;
            a.x[1] += n[1] * f;
            diff = dot2d(n, a.v);
            const double g = (BC == 5) ? (-2 * diff) : (-diff);
            a.v[0] += n[0] * g;
            a.v[1] += n[1] * g;
            set_mag(a.v, 1, a.u);
            a.phi = std::atan2(a.u[1], a.u[0]);
        }
        break;
    }
    default:
        break;
    }
}

// agents_dynamics.cpp:627-638 predictV, :642-656 predictPos
inline double predict_v(double v, double force, double t, double fr) {
    const double h = force / fr;
    return (v - h) * std::exp(-fr * t) + h;
}
inline double predict_pos(double v, double force, double t, double fr) {
    const double h = force / fr;
    return (v / fr - h / fr) * (1 - std::exp(-fr * t)) + h * t;
}
// agents_dynamics.cpp:659-696 predictXatNextBurst (the local `tc` there is computed but unused)
inline void predict_x_next_burst(const double *x, const double *v, const double *force,
                                 double t_burst, double t_tnb, double fr, double *out) {
    const double tb = std::fmin(t_burst, t_tnb);
    for (int i = 0; i < 2; i++) {
        double xa = predict_pos(v[i], force[i], tb, fr);
        const double vb = predict_v(v[i], force[i], tb, fr);
        xa += predict_pos(vb, 0., t_tnb - tb, fr);
        xa += x[i];
        out[i] = xa;
    }
}
// agents_dynamics.cpp:699-756 forceChange2NotCollide
double force_change_not_collide(const Agent &a, const sbc_params &p, double dphi) {
    const double t_burst = p.dt * p.burst_steps;
    const double t_tnb = p.dt * a.stb;
    const double fr = p.beta;
    const double fmag = len2d(a.force);
    const double fang = std::atan2(a.force[1], a.force[0]);
    double inside_change = PI;
    bool outside = true, found = false;
    double change = dphi;
    double force[2], xa[2];
    while (std::fabs(change) < PI && outside) {
        force[0] = fmag * std::cos(fang + change);
        force[1] = fmag * std::sin(fang + change);
        predict_x_next_burst(a.x, a.v, force, t_burst, t_tnb, fr, xa);
        const double rx = len2d(xa);
        if (rx < p.sizeL) {
            inside_change = change;
            found = true;
            outside = false;
        }
        if (found) dphi /= 2;
        if (outside) {
            change += dphi;
        } else if (std::fabs(dphi) > PI / 64) {
            change -= dphi;
            outside = true;
        }
    }
    return inside_change;
}
// agents_dynamics.cpp:759-772 closestForceDirection
double closest_force_direction(const Agent &a, const sbc_params &p) {
    const double fang = std::atan2(a.force[1], a.force[0]);
    const double ccw = force_change_not_collide(a, p, PI / 4);
    const double cw = force_change_not_collide(a, p, -PI / 4);
    double angle = ccw;
    if (std::fabs(cw) < std::fabs(angle)) angle = cw;
    return angle + fang;
}

enum {
    FL_BURST_START = 1, FL_SOCIAL = 2, FL_WALL = 4, FL_VPROJ_HACK = 8,
    FL_OVERSHOOT = 16, FL_BOUNDARY = 32, FL_FLEE = 64, FL_REARM = 128
};

// agents_dynamics.cpp:23-121 draw_social_or_environmental_force
void decide(Agent &a, const sbc_params &p, double u_social, double u_dir, double *force,
            double &force_mag, uint8_t &fl) {
    double hvec[2];
    force[0] = force[1] = 0;
    if (u_social <= p.prob_social) {  // :35
        fl |= FL_SOCIAL;
        force_mag = p.soc_strength;
        if (a.c_rep > 0) {  // :40 repulsion overrides
            force[0] += a.f_rep[0];
            force[1] += a.f_rep[1];
        } else {
            if (a.f_alg[0] != 0 || a.f_alg[1] != 0) {  // :47-52
                set_mag(a.f_alg, a.c_alg, hvec);
                force[0] += hvec[0];
                force[1] += hvec[1];
            }
            if (a.f_att[0] != 0 || a.f_att[1] != 0) {  // :53-58
                set_mag(a.f_att, a.c_att, hvec);
                force[0] += hvec[0];
                force[1] += hvec[1];
            }
        }
        if (force[0] != 0 || force[1] != 0) {  // :62-65
            set_mag(force, 1, hvec);
            force[0] = hvec[0];
            force[1] = hvec[1];
        } else {  // :67-71 swim straight
            force[0] = std::cos(a.phi);
            force[1] = std::sin(a.phi);
        }
    } else {
        if (a.c_flee > 0) {  // :76-80 (force_mag keeps the caller's soc_strength)
            fl |= FL_FLEE;
            set_mag(a.f_flee, 1, force);
        } else {  // :81-89
            force_mag = p.env_strength;
            const double lphi = 2 * PI * u_dir;
            force[0] = std::cos(lphi);
            force[1] = std::sin(lphi);
        }
    }
    if (p.BC >= 5) {  // :93-117 wall look-ahead
        set_mag(force, force_mag, hvec);
        force[0] = hvec[0];
        force[1] = hvec[1];
        a.force[0] = force[0];
        a.force[1] = force[1];
        double xf[2];
        predict_x_next_burst(a.x, a.v, a.force, p.dt * p.burst_steps, p.dt * a.stb, p.beta, xf);
        const double r_fut = len2d(xf);
        if (r_fut > p.sizeL - 2) {
            fl |= FL_WALL;
            const double ang = closest_force_direction(a, p);
            force[0] = force_mag * std::cos(ang);
            force[1] = force_mag * std::sin(ang);
        }
    }
    set_mag(force, force_mag, hvec);  // :119-120
    force[0] = hvec[0];
    force[1] = hvec[1];
    a.force[0] = force[0];
    a.force[1] = force[1];
}

// agents_dynamics.cpp:123-137 overshoot_check
bool overshoot(const Agent &a, const double *force, double force_mag, double lphi) {
    const double nu[2] = {std::cos(lphi), std::sin(lphi)};
    const double a0 = std::acos(dot2d(a.u, force) / force_mag);
    const double a1 = std::acos(dot2d(nu, force) / force_mag);
    const double a01 = std::acos(dot2d(nu, a.u));
    const bool o1 = (a0 < a1);
    const bool o2 = !o1 && (a01 > a0);
    const bool o3 = std::fabs(lphi - a.phi) > PI;
    return std::fabs(lphi - a.phi) > 0.01 && (o1 || o2 || o3);
}

// agents_dynamics.cpp:150-273 ParticleBurstCoast
void burst_coast(orc_world &w, Agent &a, double u_social, double u_dir, uint32_t k_next,
                 uint8_t &fl) {
    const sbc_params &p = w.p;
    double force[2] = {0, 0};
    double force_mag = p.soc_strength;
    const bool first = (a.bin_step == p.burst_steps);
    const bool bursting = (a.bin_step > 0);
    if (first) {
        fl |= FL_BURST_START;
        decide(a, p, u_social, u_dir, force, force_mag, fl);
    } else if (bursting) {
        force[0] = a.force[0];
        force[1] = a.force[1];
    } else {
        a.force[0] = force[0] = 0;
        a.force[1] = force[1] = 0;
    }
    if (bursting) a.bin_step -= 1;
    if (a.stb > 0) a.stb -= 1;

    double lphi = a.phi;
    const double vproj = a.vproj;
    const double forcev = force[0] * std::cos(lphi) + force[1] * std::sin(lphi);
    const double dt = p.dt, beta = p.beta;
    a.vproj += (-beta * vproj + forcev) * dt;  // :202
    if (a.vproj < 0) {                         // :206-210
        fl |= FL_VPROJ_HACK;
        a.vproj = 0.001;
        lphi += PI / 2;
    }
    const double forcep = -force[0] * std::sin(lphi) + force[1] * std::cos(lphi);
    lphi += p.alphaTurn * forcep * dt / vproj;  // :215 (old vproj)
    if (overshoot(a, force, force_mag, lphi)) {  // :219-223
        fl |= FL_OVERSHOOT;
        lphi = std::atan2(force[1], force[0]);
    }
    lphi = std::fmod(lphi, 2 * PI);  // :226
    a.phi = lphi;
    a.u[0] = std::cos(lphi);
    a.u[1] = std::sin(lphi);
    a.v[0] = a.vproj * a.u[0];
    a.v[1] = a.vproj * a.u[1];
    a.x[0] += a.v[0] * dt;
    a.x[1] += a.v[1] * dt;
    // :238-245
    a.f_rep[0] = a.f_rep[1] = a.f_att[0] = a.f_att[1] = 0;
    a.f_alg[0] = a.f_alg[1] = a.f_flee[0] = a.f_flee[1] = 0;
    a.c_rep = a.c_alg = a.c_att = a.c_flee = 0;
    if (p.BC != -1) {  // :247
        const double px = a.x[0], py = a.x[1];
        boundary(a, p.sizeL, p.BC);
        if ((p.BC >= 1) && (px != a.x[0] || py != a.x[1])) fl |= FL_BOUNDARY;
    }
    if (a.stb == 0) {  // :250-270 re-arm; the rejection loop is replaced by its outcome k
        fl |= FL_REARM;
        a.bin_step = p.burst_steps;
        a.stb = k_next;
    }
    if (p.BC != -1) boundary(a, p.sizeL, p.BC);  // :272
}

// agents_interact.cpp:172-250 IntCalcPrey for the directed pair (row i <- j), with
// social_forces.cpp:25-52 SFM_4Zone inlined
inline void pair_term(orc_world &w, Agent &ai, const Agent &aj, int j_id, bool gate) {
    const sbc_params &p = w.p;
    if (gate && ai.bin_step != p.burst_steps) return;  // :178
    double r[2], uji[2] = {0, 0};
    dist_vec(ai.x, aj.x, p.BC, p.sizeL, r);
    const double d = len2d(r);
    if (d > 0.0) {
        uji[0] = r[0] / d;
        uji[1] = r[1] / d;
    }
    if (d <= p.rep_range) {  // social_forces.cpp:33
        ai.f_rep[0] -= uji[0];
        ai.f_rep[1] -= uji[1];
        ai.c_rep++;
        ai.nn.push_back(j_id);
    } else if (d > p.rep_range && d <= p.alg_range) {  // :39-40
        ai.f_alg[0] += aj.v[0];
        ai.f_alg[1] += aj.v[1];
        ai.c_alg++;
        ai.nn.push_back(j_id);
    } else if (d > p.alg_range && d <= p.att_range) {  // :46-47
        ai.f_att[0] += uji[0];
        ai.f_att[1] += uji[1];
        ai.c_att++;
        ai.nn.push_back(j_id);
    }
}

// agents_interact.cpp:144-157 InteractionGlobal: every row receives its partners in ascending j
void interact_all_pairs(orc_world &w, bool gate) {
    const int N = w.N;
    for (int i = 0; i < N; i++) {
        Agent &ai = w.a[i];
        if (ai.dead) continue;
        if (gate && ai.bin_step != w.p.burst_steps) continue;
        for (int j = 0; j < N; j++) {
            if (j == i || w.a[j].dead) continue;
            pair_term(w, ai, w.a[j], j, gate);
        }
    }
}

// Delaunay edge test between alive agents i and j (Voronoi neighbour mode): some circle through both
// contains no other agent. Centre m + t n on the perpendicular bisector; every other point bounds t
// from one side; the edge exists iff max(lower) <= min(upper). Same criterion and arithmetic as the
// brute-force CGAL stand-in the reference build links against (oracle/shims/CGAL/
// Delaunay_triangulation_2.h), which is what InteractionVoronoiF2F (agents_interact.cpp:26-77) sees.
// Periodic boxes (InteractionVoronoiF2FP + GetCopies4PeriodicBC, agents_operation.cpp:240-293): the reference means to
// triangulate the agents together with shifted copies that give every agent its surroundings up to L/2 (its quadrant
// flags are never reset: undefined behaviour). Well-defined equivalent: the test for row i sees every other
// point at its image nearest to row i (CalcDistVec's minimum image, mathtools.h:54-73). In the predator phase the
// predators are vertices of the reference's triangulation (agents_interact.cpp:103-110): they can only remove edges.
inline double vor_pos(double ref, double v, int BC, double L) {
    if (BC != 0) return v;
    double r = v - ref;
    const double sign = sgn(r);
    r = std::fmod(r + sign * L / 2, L) - sign * L / 2;
    return ref + r;
}
bool delaunay_edge(const orc_world &w, int i, int j, bool with_preds) {
    const int BC = w.p.BC;
    const double L = w.p.sizeL;
    const double ax = w.a[i].x[0], ay = w.a[i].x[1];
    const double bx = vor_pos(ax, w.a[j].x[0], BC, L), by = vor_pos(ay, w.a[j].x[1], BC, L);
    const double dx = bx - ax, dy = by - ay;
    const double len2 = dx * dx + dy * dy;
    if (len2 == 0) return false;
    const double mx = 0.5 * (ax + bx), my = 0.5 * (ay + by);
    const double nx = -dy, ny = dx;
    double lo = -INFINITY, hi = INFINITY;
    const int n_pts = w.N + (with_preds ? w.Npred : 0);
    for (int k = 0; k < n_pts; k++) {
        if (k < w.N && (k == i || k == j || w.a[k].dead)) continue;
        const double *xk = (k < w.N) ? w.a[k].x : w.preds[k - w.N].x;
        const double qx = vor_pos(ax, xk[0], BC, L) - mx, qy = vor_pos(ay, xk[1], BC, L) - my;
        const double c = qx * qx + qy * qy - 0.25 * len2;
        const double s = qx * nx + qy * ny;
        if (s > 0) {
            const double b = c / (2 * s);
            if (b < hi) hi = b;
        } else if (s < 0) {
            const double b = c / (2 * s);
            if (b > lo) lo = b;
        } else if (c < 0) {
            return false;
        }
        if (lo > hi) return false;
    }
    return true;
}

// agents_interact.cpp:26-77 InteractionVoronoiF2F (non-periodic domains): both directions of every
// Delaunay edge go through IntCalcPrey; a row receives its partners in ascending id
void interact_voronoi(orc_world &w, bool gate, bool with_preds) {
    const int N = w.N;
    for (int i = 0; i < N; i++) {
        if (w.a[i].dead) continue;
        for (int j = i + 1; j < N; j++) {
            if (w.a[j].dead) continue;
            const bool need_i = !gate || w.a[i].bin_step == w.p.burst_steps;
            const bool need_j = !gate || w.a[j].bin_step == w.p.burst_steps;
            // the test runs in the frame of the receiving row (its own position, everybody else's nearest image): the
            // same edge in exact arithmetic, and in an open domain the very same floating-point decisions both ways
            if (need_i && delaunay_edge(w, i, j, with_preds)) pair_term(w, w.a[i], w.a[j], j, gate);
            if (need_j && delaunay_edge(w, j, i, with_preds)) pair_term(w, w.a[j], w.a[i], i, gate);
        }
    }
}

// with_preds: the predator phase of Voronoi mode (InteractionVoronoiF2FP, agents_interact.cpp:79-141) - the predators are
// vertices of the triangulation. Both directions of every edge are applied, as in F2F (the reference applies
// ONE direction there, which one depends on CGAL's internal edge representation).
void interact(orc_world &w, bool gate, bool with_preds = false) {
    if (w.p.neighbour_mode == SBC_NEIGHBOUR_VORONOI) interact_voronoi(w, gate, with_preds);
    else interact_all_pairs(w, gate);
}

// agents_interact.cpp:252-292 IntCalcPred
inline void detect_term(orc_world &w, Agent &ai, const Pred &pr, double u_detect) {
    const sbc_params &p = w.p;
    double r[2], uip[2] = {0, 0};
    dist_vec(ai.x, pr.x, p.BC, p.sizeL, r);
    const double ang = std::atan2(r[1], r[0]);
    const double d = len2d(r);
    if (d > 0.0) {
        uip[0] = std::cos(ang);
        uip[1] = std::sin(ang);
    }
    const double fstrength = 2 / (1 + d / p.flee_range);
    if (u_detect < fstrength) {
        ai.c_flee++;
        ai.f_flee[0] -= 1 * uip[0];
        ai.f_flee[1] -= 1 * uip[1];
    }
}

// agents_operation.cpp:23-127 GetCenterOfMass (revise=false)
void center_of_mass(orc_world &w, double *out) {
    const sbc_params &p = w.p;
    int NN = 0;
    for (const Agent &a : w.a) if (!a.dead) NN++;
    if (!p.BC) {
        double avcosx = 0, avsinx = 0, avcosy = 0, avsiny = 0;
        const double scaling = 2 * PI / p.sizeL;
        for (const Agent &a : w.a) {
            if (a.dead) continue;
            avcosx += std::cos(scaling * a.x[0]);
            avsiny += std::sin(scaling * a.x[1]);
            avcosy += std::cos(scaling * a.x[1]);
            avsinx += std::sin(scaling * a.x[0]);
        }
        avcosx /= NN; avsiny /= NN; avcosy /= NN; avsinx /= NN;
        out[0] = (std::atan2(-avsinx, -avcosx) + PI) / scaling;
        out[1] = (std::atan2(-avsiny, -avcosy) + PI) / scaling;
    } else {
        unsigned counter = 0;
        out[0] = out[1] = 0;
        for (const Agent &a : w.a) {
            if (a.dead) continue;
            counter++;
            out[0] += a.x[0];
            out[1] += a.x[1];
        }
        out[0] /= counter;
        out[1] /= counter;
    }
}

// agents_dynamics.cpp:534-555 CreatePredator
void create_predator(orc_world &w, Pred &pr, double u) {
    const sbc_params &p = w.p;
    const double phi = u * 2 * PI;
    pr.phi = phi;
    pr.u[0] = std::cos(phi);
    pr.u[1] = std::sin(phi);
    double com[2];
    center_of_mass(w, com);
    const double f = -p.sizeL / 2;
    pr.x[0] = com[0] + pr.u[0] * f;
    pr.x[1] = com[1] + pr.u[1] * f;
    pr.v[0] = p.pred_speed0 * pr.u[0];
    pr.v[1] = p.pred_speed0 * pr.u[1];
    boundary(pr, p.sizeL, p.BC);
}

// agents_dynamics.cpp:485-528 CreateFishNet
void create_fish_net(orc_world &w, double u_phi, double u_noise) {
    const sbc_params &p = w.p;
    double com[2];
    center_of_mass(w, com);
    const double net_len = p.sizeL / 4;
    double phi = 2 * PI * u_phi;
    double mv[2] = {std::cos(phi), std::sin(phi)};
    const double dist2com = 0.25;
    double start[2] = {mv[0] * (-p.sizeL * dist2com), mv[1] * (-p.sizeL * dist2com)};
    start[0] += com[0];
    start[1] += com[1];
    const double var_noise = PI / 4;
    const double noise = var_noise * u_noise - var_noise / 2;
    phi += noise;
    mv[0] = std::cos(phi);
    mv[1] = std::sin(phi);
    const double perp[2] = {mv[1], -mv[0]};  // mathtools.h:183-189 vec_perp
    start[0] += perp[0] * (-net_len / 2);
    start[1] += perp[1] * (-net_len / 2);
    const double net_dist = net_len / (w.preds.size() - 1);
    for (unsigned i = 0; i < w.preds.size(); i++) {
        Pred &pr = w.preds[i];
        pr.x[0] = start[0] + perp[0] * (i * net_dist);
        pr.x[1] = start[1] + perp[1] * (i * net_dist);
        pr.phi = phi;
        pr.u[0] = mv[0];
        pr.u[1] = mv[1];
        pr.v[0] = p.pred_speed0 * pr.u[0];
        pr.v[1] = p.pred_speed0 * pr.u[1];
        boundary(pr, p.sizeL, p.BC);
    }
}

// agents_dynamics.cpp:567-613 MovePredator
void move_predator(orc_world &w, Pred &pr, double gauss) {
    const sbc_params &p = w.p;
    double lphi;
    if (p.pred_move == 0) {
        const double rnp = p.noisep * gauss;
        lphi = pr.phi;
        lphi += rnp * std::sqrt(p.pred_speed0);
        lphi = std::fmod(lphi, 2 * PI);
    } else {
        double minvec[2] = {0, 0}, hv[2];
        double mindist = 0;
        bool have = false;
        for (const Agent &a : w.a) {
            if (a.dead) continue;
            dist_vec(pr.x, a.x, p.BC, p.sizeL, hv);
            const double d = len2d(hv);
            if (!have || d < mindist) {
                have = true;
                mindist = d;
                minvec[0] = hv[0];
                minvec[1] = hv[1];
            }
        }
        lphi = std::atan2(minvec[1], minvec[0]);
    }
    pr.phi = lphi;
    pr.u[0] = std::cos(lphi);
    pr.u[1] = std::sin(lphi);
    pr.v[0] = p.pred_speed0 * pr.u[0];
    pr.v[1] = p.pred_speed0 * pr.u[1];
    pr.x[0] = pr.x[0] + pr.v[0] * p.dt;
    pr.x[1] = pr.x[1] + pr.v[1] * p.dt;
    boundary(pr, p.sizeL, p.BC);
}

// agents_dynamics.cpp:780-820 FishNetKill
void fish_net_kill(orc_world &w) {
    const sbc_params &p = w.p;
    double vnet[2];
    dist_vec(w.preds[0].x, w.preds[1].x, p.BC, p.sizeL, vnet);
    const double dist = len2d(vnet);
    const double net_len = (w.preds.size() - 1) * dist;
    // note: the reference calls vec_div(v_net, dist) and DISCARDS the result (:787), so the
    // un-normalised v_net is what `side` is projected on
    const double *mv = w.preds[0].u;
    double r[2];
    for (Agent &a : w.a) {
        if (a.dead) continue;
        dist_vec(w.preds[0].x, a.x, p.BC, p.sizeL, r);
        const double front = dot2d(mv, r);
        const double side = dot2d(vnet, r);
        if (front > 0 && front < p.kill_range && side > 0 && side <= net_len) {
            a.dead = true;
            w.n_dead++;
        }
    }
}

// mathtools.cpp:22-33 SigmoidFunction
inline double sigmoid(double x, double x0, double steep) {
    return 0.5 * (std::tanh(steep * (x - x0)) + 1.0);
}

// agents_dynamics.cpp:826-915 PredKill; `luck` of candidate i is the agent's own draw
void pred_kill(orc_world &w, const Pred &pr, int64_t step) {
    const sbc_params &p = w.p;
    const double r_sense = 4 * p.kill_range;
    double r[2];
    if (p.pred_kill == 1) {
        for (Agent &a : w.a) {
            if (a.dead) continue;
            dist_vec(pr.x, a.x, p.BC, p.sizeL, r);
            if (len2d(r) <= p.kill_range) {
                a.dead = true;
                w.n_dead++;
            }
        }
    } else if (p.pred_kill > 1) {
        std::vector<int> cand;
        std::vector<double> probs;
        int n_sensed = 0;
        double total = 0;
        for (int i = 0; i < w.N; i++) {
            Agent &a = w.a[i];
            if (a.dead) continue;
            dist_vec(pr.x, a.x, p.BC, p.sizeL, r);
            const double d = len2d(r);
            if (d <= r_sense) {
                n_sensed++;
                if (d < p.kill_range) {
                    cand.push_back(i);
                    const double pc = (p.kill_range - d) / p.kill_range;
                    probs.push_back(pc);
                    total += pc;
                }
            }
        }
        double eff = 0;
        if (n_sensed > 0) eff = p.kill_rate * sigmoid(n_sensed, p.N_confu, -1);
        for (size_t c = 0; c < cand.size(); c++) {
            double pk = 0;
            if (p.pred_kill == 2) pk = p.kill_rate * p.dt;
            else if (p.pred_kill == 3) pk = eff * p.dt;
            else if (p.pred_kill == 4) pk = eff * p.dt * (probs[c] / total);
            const OrcPhilox4 q = orc_draw(w.seed, w.replicate, (uint32_t)cand[c], step, ORC_SLOT_DECISION);
            const double luck = orc_u32_to_unit(q.v[3]);
            if (pk > luck) {
                w.a[cand[c]].dead = true;
                w.n_dead++;
            }
        }
    }
}

double gauss_from(uint32_t r0, uint32_t r1) {
    // Box-Muller on two 32-bit draws (protocol)
    const double u1 = (r0 + 1.0) * (1.0 / 4294967296.0);
    const double u2 = orc_u32_to_unit(r1);
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * PI * u2);
}

// swarmdyn.cpp:143-204 Step, metric neighbour mode
void one_step(orc_world &w, int64_t s, uint8_t *flags) {
    const sbc_params &p = w.p;
    const double dt = p.dt;
    const int N = w.N;
    const bool have_pred = w.Npred > 0;
    // :153-158 spawn
    if (have_pred && s >= p.pred_time / dt && s - 1 < p.pred_time / dt) {
        const OrcPhilox4 q = orc_draw(w.seed, w.replicate, ORC_PREDATOR_AGENT_ID, s, 0);
        if (w.Npred == 1) create_predator(w, w.preds[0], orc_u32_to_unit(q.v[0]));
        else create_fish_net(w, orc_u32_to_unit(q.v[0]), orc_u32_to_unit(q.v[1]));
    }
    // :160-170 net re-spawn (creations before pred_time are unobservable and skipped)
    if (w.Npred > 1 && s >= p.pred_time / dt) {
        const int since = (int)s - (int)(p.pred_time / dt);
        int needed;
        if (p.pred_speed0 > 0) needed = (int)(p.sizeL / 2 / (p.pred_speed0 * dt));
        else needed = (int)((p.sim_time - p.trans_time) / 4 / dt);
        if (since % needed == 0) {
            const OrcPhilox4 q = orc_draw(w.seed, w.replicate, ORC_PREDATOR_AGENT_ID, s, 1);
            create_fish_net(w, orc_u32_to_unit(q.v[0]), orc_u32_to_unit(q.v[1]));
        }
    }
    for (Agent &a : w.a) a.nn.clear();  // :172-173
    // :179-182 interaction
    interact(w, true, have_pred && !(s < p.pred_time / dt));
    if (have_pred && !(s < p.pred_time / dt)) {
        for (int i = 0; i < N; i++) {
            Agent &a = w.a[i];
            if (a.dead) continue;
            for (int j = 0; j < w.Npred; j++) {
                const OrcPhilox4 q = orc_draw(w.seed, w.replicate, (uint32_t)i, s,
                                              ORC_SLOT_DETECT0 + (uint32_t)(j / 4));
                detect_term(w, a, w.preds[j], orc_u32_to_unit(q.v[j % 4]));
            }
        }
    }
    // :185-190 update
    for (int i = 0; i < N; i++) {
        Agent &a = w.a[i];
        if (a.dead) continue;
        const OrcPhilox4 q = orc_draw(w.seed, w.replicate, (uint32_t)i, s, ORC_SLOT_DECISION);
        const double us = w.inj_social ? w.u_social[i] : orc_u32_to_unit(q.v[0]);
        const double ud = w.inj_dir ? w.u_dir[i] : orc_u24_to_unit(q.v[1]);
        const uint32_t kn = w.inj_k ? w.k_next[i] : orc_geometric_k(w.geo, q.v[2]);
        uint8_t fl = 0;
        burst_coast(w, a, us, ud, kn, fl);
        if (flags) flags[i] |= fl;
    }
    // :192-203 predator
    if (have_pred && s >= p.pred_time / dt) {
        if (w.Npred > 1) {
            for (Pred &pr : w.preds) {  // agents_dynamics.cpp:616-624 MoveFishNet
                pr.x[0] = pr.x[0] + pr.v[0] * dt;
                pr.x[1] = pr.x[1] + pr.v[1] * dt;
                boundary(pr, p.sizeL, p.BC);
            }
            if (p.pred_kill != 0) fish_net_kill(w);
        } else {
            const OrcPhilox4 q = orc_draw(w.seed, w.replicate, ORC_PREDATOR_AGENT_ID, s, 2);
            move_predator(w, w.preds[0], gauss_from(q.v[0], q.v[1]));
            if (p.pred_kill != 0) pred_kill(w, w.preds[0], s);
        }
    }
    w.inj_social = w.inj_dir = w.inj_k = false;
}

}  // namespace

extern "C" {

const char *orc_kind(void) { return "port"; }

orc_world *orc_create(const sbc_params *p, int n_agents, int n_pred, uint64_t seed,
                      uint64_t replicate_id) {
    orc_world *w = new orc_world();
    w->p = *p;
    w->N = n_agents;
    w->Npred = n_pred;
    w->seed = seed;
    w->replicate = replicate_id;
    w->n_dead = 0;
    w->inj_social = w->inj_dir = w->inj_k = false;
    w->a.resize(n_agents);
    for (Agent &a : w->a) {  // settings.cpp:168-193 InitSystem defaults
        std::memset(a.x, 0, sizeof(double) * 2);
        a.v[0] = a.v[1] = a.u[0] = a.u[1] = 0;
        a.phi = 0;
        a.vproj = 1;
        a.force[0] = a.force[1] = 0;
        a.f_rep[0] = a.f_rep[1] = a.f_alg[0] = a.f_alg[1] = 0;
        a.f_att[0] = a.f_att[1] = a.f_flee[0] = a.f_flee[1] = 0;
        a.c_rep = a.c_alg = a.c_att = a.c_flee = 0;
        a.bin_step = a.stb = 0;
        a.dead = false;
    }
    w->preds.resize(n_pred);
    for (Pred &pr : w->preds) {  // settings.cpp:147-166 InitPredator
        pr.x[0] = pr.x[1] = pr.v[0] = pr.v[1] = pr.u[0] = pr.u[1] = 1.;
        pr.phi = 1;
    }
    w->geo = orc_geometric_table(p->burst_rate, p->dt);
    return w;
}
void orc_destroy(orc_world *w) { delete w; }

void orc_set_state(orc_world *w, const double *x, const double *y, const double *vx,
                   const double *vy, const double *phi, const double *vproj, const double *fx,
                   const double *fy, const uint32_t *bin_step, const uint32_t *stb,
                   const uint8_t *alive) {
    for (int i = 0; i < w->N; i++) {
        Agent &a = w->a[i];
        if (x) a.x[0] = x[i];
        if (y) a.x[1] = y[i];
        if (phi) {
            a.phi = phi[i];
            a.u[0] = std::cos(a.phi);
            a.u[1] = std::sin(a.phi);
        }
        if (vproj) a.vproj = vproj[i];
        if (vx) a.v[0] = vx[i]; else if (phi) a.v[0] = a.vproj * a.u[0];
        if (vy) a.v[1] = vy[i]; else if (phi) a.v[1] = a.vproj * a.u[1];
        if (fx) a.force[0] = fx[i];
        if (fy) a.force[1] = fy[i];
        if (bin_step) a.bin_step = bin_step[i];
        if (stb) a.stb = stb[i];
        if (alive) a.dead = !alive[i];
    }
    w->n_dead = 0;
    for (const Agent &a : w->a) w->n_dead += a.dead;
}
void orc_get_state(orc_world *w, double *x, double *y, double *vx, double *vy, double *phi,
                   double *vproj, double *fx, double *fy, uint32_t *bin_step, uint32_t *stb,
                   uint8_t *alive) {
    for (int i = 0; i < w->N; i++) {
        const Agent &a = w->a[i];
        if (x) x[i] = a.x[0];
        if (y) y[i] = a.x[1];
        if (vx) vx[i] = a.v[0];
        if (vy) vy[i] = a.v[1];
        if (phi) phi[i] = a.phi;
        if (vproj) vproj[i] = a.vproj;
        if (fx) fx[i] = a.force[0];
        if (fy) fy[i] = a.force[1];
        if (bin_step) bin_step[i] = a.bin_step;
        if (stb) stb[i] = a.stb;
        if (alive) alive[i] = !a.dead;
    }
}
void orc_set_predators(orc_world *w, const double *x, const double *y, const double *phi, int) {
    for (int j = 0; j < w->Npred; j++) {
        Pred &pr = w->preds[j];
        pr.x[0] = x[j];
        pr.x[1] = y[j];
        pr.phi = phi[j];
        pr.u[0] = std::cos(phi[j]);
        pr.u[1] = std::sin(phi[j]);
        pr.v[0] = w->p.pred_speed0 * pr.u[0];
        pr.v[1] = w->p.pred_speed0 * pr.u[1];
    }
}
void orc_get_predators(orc_world *w, double *out) {
    for (int j = 0; j < w->Npred; j++) {
        const Pred &pr = w->preds[j];
        out[6 * j + 0] = pr.x[0];
        out[6 * j + 1] = pr.x[1];
        out[6 * j + 2] = pr.v[0];
        out[6 * j + 3] = pr.v[1];
        out[6 * j + 4] = 0;
        out[6 * j + 5] = 0;
    }
}
void orc_get_counts(orc_world *w, int32_t *n_alive, int32_t *n_dead) {
    int d = 0;
    for (const Agent &a : w->a) d += a.dead;
    if (n_alive) *n_alive = w->N - d;
    if (n_dead) *n_dead = d;
}

int64_t orc_pair_interactions(orc_world *w, int flags, int32_t *c_rep, int32_t *c_alg,
                              int32_t *c_att, double *f_rep, double *f_alg, double *f_att,
                              int64_t *nn_ptr, int32_t *nn_idx, int64_t nn_capacity) {
    for (Agent &a : w->a) {
        a.f_rep[0] = a.f_rep[1] = a.f_alg[0] = a.f_alg[1] = a.f_att[0] = a.f_att[1] = 0;
        a.c_rep = a.c_alg = a.c_att = 0;
        a.nn.clear();
    }
    interact(*w, flags == SBC_PAIR_ACTIVE_ROWS);
    int64_t total = 0;
    for (int i = 0; i < w->N; i++) {
        Agent &a = w->a[i];
        if (c_rep) c_rep[i] = a.c_rep;
        if (c_alg) c_alg[i] = a.c_alg;
        if (c_att) c_att[i] = a.c_att;
        if (f_rep) { f_rep[2 * i] = a.f_rep[0]; f_rep[2 * i + 1] = a.f_rep[1]; }
        if (f_alg) { f_alg[2 * i] = a.f_alg[0]; f_alg[2 * i + 1] = a.f_alg[1]; }
        if (f_att) { f_att[2 * i] = a.f_att[0]; f_att[2 * i + 1] = a.f_att[1]; }
        if (nn_ptr) nn_ptr[i] = total;
        std::sort(a.nn.begin(), a.nn.end());
        if (nn_idx)
            for (size_t k = 0; k < a.nn.size(); k++)
                if (total + (int64_t)k < nn_capacity) nn_idx[total + k] = a.nn[k];
        total += (int64_t)a.nn.size();
    }
    if (nn_ptr) nn_ptr[w->N] = total;
    for (Agent &a : w->a) {  // leave the accumulators as a fresh step expects them (:238-245)
        a.f_rep[0] = a.f_rep[1] = a.f_alg[0] = a.f_alg[1] = a.f_att[0] = a.f_att[1] = 0;
        a.c_rep = a.c_alg = a.c_att = 0;
        a.nn.clear();
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
        int alive = 0;
        for (const Agent &a : w->a) alive += !a.dead;
        if (alive == 0) break;  // swarmdyn.cpp:72-75
        one_step(*w, s, branch_flags);
    }
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
    // basic_particle_averages swarmdyn.cpp:486-510 and the position mean / polarisation of
    // Out_swarm :258-298; NND here is the true nearest-neighbour distance (the reference's
    // loop skips neighbour i-1, :275)
    const sbc_params &p = w->p;
    int n = 0;
    double ax[2] = {0, 0}, av[2] = {0, 0}, au[2] = {0, 0}, as = 0, nnd = 0;
    for (int i = 0; i < w->N; i++) {
        const Agent &a = w->a[i];
        if (a.dead) continue;
        n++;
        ax[0] += a.x[0]; ax[1] += a.x[1];
        av[0] += a.v[0]; av[1] += a.v[1];
        au[0] += a.u[0]; au[1] += a.u[1];
        as += std::sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1]);
        double best = -1, r[2];
        for (int j = 0; j < w->N; j++) {
            if (j == i || w->a[j].dead) continue;
            dist_vec(a.x, w->a[j].x, p.BC, p.sizeL, r);
            const double d = len2d(r);
            if (best < 0 || d < best) best = d;
        }
        if (best >= 0) nnd += best;
    }
    out[0] = n;
    if (n == 0) { for (int k = 1; k < 8; k++) out[k] = 0; return; }
    au[0] /= n; au[1] /= n;
    out[1] = len2d(au);
    out[2] = ax[0] / n; out[3] = ax[1] / n;
    out[4] = av[0] / n; out[5] = av[1] / n;
    out[6] = as / n;
    out[7] = nnd / n;
}

// Out_swarm, swarmdyn.cpp:237-344 (with basic_particle_averages :486-510, get_elongation agents_operation.cpp:153-177,
// AreaConvexHull :130-150 as a monotone-chain hull): the 16 columns of /swarm, quirks included - the nearest-neighbour
// loop runs j < i-1 (:275) from N * alg_range (:273), IID sums j > i and is divided by N(N-1) (:296), ND divides by the
// size of a possibly empty neighbour list (:270). Agents = the alive ones in id order (split_dead, :71).
namespace {
double port_elongation(const std::vector<const Agent *> &a, const double *dir) {
    const double len = len2d(dir);
    const double c[2] = {dir[0] / len, dir[1] / len}, q[2] = {c[1], -c[0]};
    double mn = std::numeric_limits<double>::max(), mx = std::numeric_limits<double>::lowest(), pmn = mn, pmx = mx;
    for (const Agent *g : a) {
        const double d = g->x[0] * c[0] + g->x[1] * c[1], pd = g->x[0] * q[0] + g->x[1] * q[1];
        mn = std::fmin(mn, d); mx = std::fmax(mx, d); pmn = std::fmin(pmn, pd); pmx = std::fmax(pmx, pd);
    }
    return std::fabs((mx - mn) / (pmx - pmn));
}
double port_hull_area(const std::vector<const Agent *> &a) {
    std::vector<std::pair<double, double>> pts;
    for (const Agent *g : a) pts.push_back({g->x[0], g->x[1]});
    std::sort(pts.begin(), pts.end());
    const int n = (int)pts.size();
    if (n < 3) return 0;
    std::vector<std::pair<double, double>> h(2 * n);
    auto cross = [](const std::pair<double, double> &o, const std::pair<double, double> &p, const std::pair<double, double> &q) {
        return (p.first - o.first) * (q.second - o.second) - (p.second - o.second) * (q.first - o.first);
    };
    int k = 0;
    for (int i = 0; i < n; i++) { while (k >= 2 && cross(h[k - 2], h[k - 1], pts[i]) <= 0) k--; h[k++] = pts[i]; }
    for (int i = n - 2, t = k + 1; i >= 0; i--) { while (k >= t && cross(h[k - 2], h[k - 1], pts[i]) <= 0) k--; h[k++] = pts[i]; }
    double area = 0;
    for (int i = 0; i + 1 < k; i++) area += h[i].first * h[i + 1].second - h[i + 1].first * h[i].second;
    return 0.5 * area;
}
}  // namespace

void orc_out_swarm(orc_world *w, double *out16) {
    const sbc_params &p = w->p;
    std::vector<const Agent *> a;
    for (const Agent &g : w->a) if (!g.dead) a.push_back(&g);
    const int N = (int)a.size();
    double avg_v[2] = {0, 0}, avg_u[2] = {0, 0}, avg_s = 0, avg_v2 = 0, avg_x[2] = {0, 0};
    for (const Agent *g : a) {   // :486-510
        avg_v[0] += g->v[0]; avg_v[1] += g->v[1]; avg_u[0] += g->u[0]; avg_u[1] += g->u[1];
        const double v2 = g->v[0] * g->v[0] + g->v[1] * g->v[1];
        avg_s += std::sqrt(v2); avg_v2 += v2;
    }
    for (int d = 0; d < 2; d++) { avg_v[d] /= N; avg_u[d] /= N; }
    avg_s /= N; avg_v2 /= N;
    double IID = 0, NND = 0, ND = 0, max_IID = 0, max_vec[2] = {1, 1}, hd = 0, r[2];
    for (int i = 0; i < N; i++) {
        avg_x[0] += a[i]->x[0]; avg_x[1] += a[i]->x[1];
        double nd = 0;
        for (int j : a[i]->nn) { dist_vec(a[i]->x, w->a[j].x, p.BC, p.sizeL, r); nd += len2d(r); }   // :263-271
        nd /= (double)a[i]->nn.size();
        ND += nd;
        hd = w->N * p.alg_range;                       // :273
        for (int j = 0; j < i - 1; j++) { dist_vec(a[i]->x, a[j]->x, p.BC, p.sizeL, r); hd = std::fmin(hd, len2d(r)); }   // :275
        for (int j = i + 1; j < N; j++) {              // :280-289
            dist_vec(a[i]->x, a[j]->x, p.BC, p.sizeL, r);
            const double d = len2d(r);
            hd = std::fmin(hd, d);
            IID += d;
            if (d > max_IID) { max_IID = d; max_vec[0] = r[0]; max_vec[1] = r[1]; }
        }
        NND += hd;
    }
    avg_x[0] /= N; avg_x[1] /= N;
    NND /= N; ND /= N;
    IID /= (double)((N - 1) * N);
    const double avgvel = len2d(avg_v), pol = len2d(avg_u);
    double L_norm = 0;
    for (const Agent *g : a) {   // :301-306
        const double h[2] = {g->x[0] - avg_x[0], g->x[1] - avg_x[1]};
        L_norm += (h[0] * g->v[1] - h[1] * g->v[0]) / len2d(h);
    }
    L_norm = std::fabs(L_norm) / N;
    const double area = port_hull_area(a), elong = port_elongation(a, avg_u), aspect = port_elongation(a, max_vec);
    double a1 = (avg_v[0] * max_vec[0] + avg_v[1] * max_vec[1]) / (avgvel * len2d(max_vec));
    const double a2 = std::acos(-a1);
    a1 = std::acos(a1);
    const double o[16] = {(double)N, pol, L_norm, avg_x[0], avg_x[1], avg_v[0], avg_v[1], avg_s, avg_v2, elong, aspect,
                          std::fmin(a1, a2), area, NND, IID, ND};
    for (int k = 0; k < 16; k++) out16[k] = o[k];
}

// Out_swarm_fishNet, swarmdyn.cpp:358-410
void orc_out_swarm_fishnet(orc_world *w, double *out4) {
    const sbc_params &p = w->p;
    int n_front_close = 0, n_front = 0, n = 0;
    double r[2];
    if (w->Npred > 1) {
        const double *mv = w->preds[0].u;
        double vnet[2];
        dist_vec(w->preds[0].x, w->preds[1].x, p.BC, p.sizeL, vnet);
        const double net_len = (w->Npred - 1) * len2d(vnet);
        for (const Agent &g : w->a) {
            if (g.dead) continue;
            dist_vec(w->preds[0].x, g.x, p.BC, p.sizeL, r);
            const double front = mv[0] * r[0] + mv[1] * r[1], side = vnet[0] * r[0] + vnet[1] * r[1];
            if (side > 0 && side <= net_len) { n_front++; if (front > 0 && front < p.kill_range) n_front_close++; }
        }
    }
    double dist_net = 0;
    for (const Agent &g : w->a) {
        if (g.dead) continue;
        n++;
        double dmin = p.sizeL * 2;
        for (int j = 0; j < w->Npred; j++) { dist_vec(w->preds[j].x, g.x, p.BC, p.sizeL, r); dmin = std::fmin(dmin, len2d(r)); }
        dist_net += dmin;
    }
    dist_net /= n;
    int dead = 0;
    for (const Agent &g : w->a) dead += g.dead;
    out4[0] = n_front_close; out4[1] = dist_net; out4[2] = n_front; out4[3] = dead;
}

// ResetSystem, settings.cpp:196-387, on the repo's IC protocol (orc_philox.h): cases 0-5 and the "else" branch, unit
// speed (:368-378), Boundary (:384-386). The ring shuffle of IC 3-5 is the protocol's keyed bijection (the reference
// shuffles with a clock seed). Returns -1 for IC 99 (a file).
int orc_init_state(orc_world *w, int ic) {
    if (ic == 99) return -1;
    const sbc_params &p = w->p;
    const int N = w->N;
    const double L = p.sizeL, range = 3 * p.rep_range;
    const double theta_all = 2 * M_PI * orc_u32_to_unit(orc_draw(w->seed, w->replicate, ORC_IC_SWARM_ID, ORC_IC_STEP, 16).v[0]);
    const OrcPhilox4 key = orc_draw(w->seed, w->replicate, ORC_IC_SWARM_ID, ORC_IC_STEP, 17);
    int circle = (ic == 3) ? 1 : 3, Ncir = (ic == 3) ? 3 : int(2 * M_PI * circle), Ncap = Ncir, ccircle = 0;
    const double sigma_space = (ic == 3) ? 0.8 : 1;
    for (int i = 0; i < N; i++) {
        const OrcPhilox4 d = orc_draw(w->seed, w->replicate, (uint32_t)i, ORC_IC_STEP, 16);
        const double ux = orc_u32_to_unit(d.v[0]), uy = orc_u32_to_unit(d.v[1]), up = orc_u32_to_unit(d.v[2]);
        int ii = i;
        double x, y, phi;
        if (ic == 1) { x = L * ux; y = L * uy; phi = theta_all; }                                  // :212-221
        else if (ic == 0 || ic == 2) {                                                             // :237-257
            x = std::fmin(L, std::sqrt(N) * range) * ux;
            y = std::fmin(L, std::sqrt(N) * range) * uy;
            phi = (ic == 0) ? 2 * M_PI * up : theta_all;
        } else if (ic == 3 || ic == 4 || ic == 5) {                                                // :258-355
            ii = (int)orc_ic_permute((uint32_t)i, (uint32_t)N, key);
            if (i == Ncap) { circle += 1; Ncir = int(2 * M_PI * circle); Ncap += Ncir; ccircle = 0; }
            const double angle = 2 * M_PI * ccircle / Ncir;
            x = range * circle * std::cos(angle) + sigma_space * range * ux;
            y = range * circle * std::sin(angle) + sigma_space * range * uy;
            phi = (ic == 3) ? theta_all : (ic == 4) ? std::fmod(angle + M_PI / 2., 2 * M_PI) : 2 * M_PI * up;
            ccircle += 1;
        } else { x = L * ux; y = L * uy; phi = 2 * M_PI * up; }                                    // :357-366
        Agent &a = w->a[ii];
        a.x[0] = x; a.x[1] = y; a.phi = phi;
        a.v[0] = a.u[0] = std::cos(phi); a.v[1] = a.u[1] = std::sin(phi);                          // :368-378
        a.vproj = 1;
        a.force[0] = a.force[1] = 0;
        a.bin_step = a.stb = 0;
        a.dead = false;
        if (p.BC != -1) boundary(a, L, p.BC);   // :384-386; phi follows v only at the circular walls, as in the reference
    }
    w->n_dead = 0;
    return 0;
}

int orc_ref_voronoi_step(orc_world *, int64_t, int64_t, uint32_t) { return -1; }

double orc_timed_steps(orc_world *w, int64_t s_first, int64_t n_steps) {
    const auto t0 = std::chrono::steady_clock::now();
    orc_step(w, s_first, n_steps, nullptr);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
