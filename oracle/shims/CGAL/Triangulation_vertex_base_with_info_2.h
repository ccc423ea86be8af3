// Oracle shim: vertex type carrying an info() payload.
#ifndef SBC_ORACLE_SHIM_CGAL_TVBWI_H
#define SBC_ORACLE_SHIM_CGAL_TVBWI_H
#include "Exact_predicates_inexact_constructions_kernel.h"
namespace CGAL {
template <class Info, class K>
struct Triangulation_vertex_base_with_info_2 {
    typedef Info Info_type;
    typename K::Point_2 pt;
    Info inf;
    Info &info() { return inf; }
    const typename K::Point_2 &point() const { return pt; }
};
template <class Vb>
struct Triangulation_data_structure_2 {
    typedef Vb Vertex;
};
}
#endif
