// Oracle shim: included by the reference, nothing used from it on the hot path.
#ifndef SBC_ORACLE_SHIM_CGAL_Alpha_shape_2_H
#define SBC_ORACLE_SHIM_CGAL_Alpha_shape_2_H
#endif
