// Oracle shim: polygon with signed shoelace area (agents_operation.cpp:143-147).
#ifndef SBC_ORACLE_SHIM_CGAL_POLYGON_H
#define SBC_ORACLE_SHIM_CGAL_POLYGON_H
#include <vector>
#include "Exact_predicates_inexact_constructions_kernel.h"
namespace CGAL {
template <class K>
class Polygon_2 {
    std::vector<typename K::Point_2> p;
public:
    void push_back(const typename K::Point_2 &q) { p.push_back(q); }
    double area() const {
        double a = 0;
        const size_t n = p.size();
        for (size_t i = 0; i < n; i++) {
            const size_t j = (i + 1) % n;
            a += p[i].x() * p[j].y() - p[j].x() * p[i].y();
        }
        return 0.5 * a;
    }
};
}
#endif
