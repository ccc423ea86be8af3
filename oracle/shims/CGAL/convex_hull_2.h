// Oracle shim: Andrew monotone-chain convex hull, counter-clockwise, collinear points dropped
// (same output convention as CGAL::convex_hull_2; used at agents_operation.cpp:140).
#ifndef SBC_ORACLE_SHIM_CGAL_CONVEX_HULL_H
#define SBC_ORACLE_SHIM_CGAL_CONVEX_HULL_H
#include <algorithm>
#include <vector>
#include "Exact_predicates_inexact_constructions_kernel.h"
namespace CGAL {
template <class InIt, class OutIt>
OutIt convex_hull_2(InIt first, InIt last, OutIt out) {
    typedef Shim_point_2 P;
    std::vector<P> pts(first, last);
    std::sort(pts.begin(), pts.end(), [](const P &a, const P &b) {
        return a.x() < b.x() || (a.x() == b.x() && a.y() < b.y());
    });
    pts.erase(std::unique(pts.begin(), pts.end(),
                          [](const P &a, const P &b) { return a.x() == b.x() && a.y() == b.y(); }),
              pts.end());
    const size_t n = pts.size();
    if (n < 3) {
        for (size_t i = 0; i < n; i++) *out++ = pts[i];
        return out;
    }
    auto cross = [](const P &o, const P &a, const P &b) {
        return (a.x() - o.x()) * (b.y() - o.y()) - (a.y() - o.y()) * (b.x() - o.x());
    };
    std::vector<P> h(2 * n);
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        while (k >= 2 && cross(h[k - 2], h[k - 1], pts[i]) <= 0) k--;
        h[k++] = pts[i];
    }
    for (size_t i = n - 1, t = k + 1; i > 0; i--) {
        while (k >= t && cross(h[k - 2], h[k - 1], pts[i - 1]) <= 0) k--;
        h[k++] = pts[i - 1];
    }
    for (size_t i = 0; i + 1 < k; i++) *out++ = h[i];
    return out;
}
}
#endif
