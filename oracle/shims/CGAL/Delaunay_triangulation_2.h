// Oracle shim (test infrastructure only): brute-force 2-D Delaunay EDGE enumeration exposing the
// sliver of the CGAL API the reference touches (agents_interact.cpp:29-36,62-73).
//
// (i,j) is a Delaunay edge iff some circle through p_i and p_j contains no other point. Put the
// centre at m + t*n on the perpendicular bisector of ij: every other point k confines t to a
// half-line, so the edge exists iff max(lower bounds) <= min(upper bounds). O(n) per pair,
// O(n^3) overall - fine for oracle-sized inputs. Exactly co-circular ties keep both diagonals
// (CGAL would keep one); such inputs have measure zero for the random states used in tests.
#ifndef SBC_ORACLE_SHIM_CGAL_DELAUNAY_H
#define SBC_ORACLE_SHIM_CGAL_DELAUNAY_H
#include <limits>
#include <utility>
#include <vector>
#include "Triangulation_vertex_base_with_info_2.h"
namespace CGAL {
template <class K, class Tds>
class Delaunay_triangulation_2 {
public:
    typedef typename K::Point_2 Point;
    typedef typename Tds::Vertex Vertex;
    typedef Vertex *Vertex_handle;
    struct Face {
        Vertex_handle v[3];
        Vertex_handle vertex(int i) const { return v[i]; }
        static int ccw(int i) { return (i + 1) % 3; }
        static int cw(int i) { return (i + 2) % 3; }
    };
    typedef Face *Face_handle;
    typedef std::pair<Face_handle, int> Edge;
    typedef typename std::vector<Edge>::iterator Edge_iterator;

    template <class It>
    void insert(It first, It last) {
        for (It it = first; it != last; ++it) {
            Vertex v;
            v.pt = it->first;
            v.inf = it->second;
            verts.push_back(v);
        }
        build();
    }
    Edge_iterator finite_edges_begin() { return edges.begin(); }
    Edge_iterator finite_edges_end() { return edges.end(); }

private:
    std::vector<Vertex> verts;
    std::vector<Face> faces;
    std::vector<Edge> edges;

    void build() {
        const size_t n = verts.size();
        faces.clear();
        edges.clear();
        faces.reserve(n * (n - 1) / 2 + 1);
        for (size_t i = 0; i < n; i++) {
            for (size_t j = i + 1; j < n; j++) {
                const double ax = verts[i].pt.x(), ay = verts[i].pt.y();
                const double bx = verts[j].pt.x(), by = verts[j].pt.y();
                const double dx = bx - ax, dy = by - ay;
                const double len2 = dx * dx + dy * dy;
                if (len2 == 0) continue;
                const double mx = 0.5 * (ax + bx), my = 0.5 * (ay + by);
                const double nx = -dy, ny = dx;  // left normal of i->j
                double lo = -std::numeric_limits<double>::infinity();
                double hi = std::numeric_limits<double>::infinity();
                bool ok = true;
                for (size_t k = 0; k < n && ok; k++) {
                    if (k == i || k == j) continue;
                    const double qx = verts[k].pt.x() - mx, qy = verts[k].pt.y() - my;
                    // k outside-or-on circle centred m+t*n through a:  |q - t n|^2 >= |a-m|^2 + t^2 |n|^2
                    //   <=>  |q|^2 - len2/4 >= 2 t (q.n)
                    const double c = qx * qx + qy * qy - 0.25 * len2;
                    const double s = qx * nx + qy * ny;
                    if (s > 0) {
                        const double b = c / (2 * s);
                        if (b < hi) hi = b;
                    } else if (s < 0) {
                        const double b = c / (2 * s);
                        if (b > lo) lo = b;
                    } else if (c < 0) {
                        ok = false;  // collinear point strictly between i and j
                    }
                    if (lo > hi) ok = false;
                }
                if (!ok) continue;
                Face f;
                f.v[0] = nullptr;
                f.v[1] = &verts[j];  // ccw(0)
                f.v[2] = &verts[i];  // cw(0)
                faces.push_back(f);
            }
        }
        edges.reserve(faces.size());
        for (size_t e = 0; e < faces.size(); e++) edges.push_back(Edge(&faces[e], 0));
    }
};
}
#endif
