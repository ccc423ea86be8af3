// Oracle shim (test infrastructure only): minimal 2-D kernel with the Point_2 type the reference
// constructs (agents_interact.cpp:45, agents_operation.cpp:131-138).
#ifndef SBC_ORACLE_SHIM_CGAL_EPICK_H
#define SBC_ORACLE_SHIM_CGAL_EPICK_H
namespace CGAL {
struct Shim_point_2 {
    double px, py;
    Shim_point_2() : px(0), py(0) {}
    Shim_point_2(double x_, double y_) : px(x_), py(y_) {}
    double x() const { return px; }
    double y() const { return py; }
};
struct Exact_predicates_inexact_constructions_kernel {
    typedef Shim_point_2 Point_2;
};
}
#endif
