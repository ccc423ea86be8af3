// Oracle shim: included by the reference, nothing used from it on the hot path.
#ifndef SBC_ORACLE_SHIM_CGAL_algorithm_H
#define SBC_ORACLE_SHIM_CGAL_algorithm_H
#endif
