// Oracle shim: the reference includes this header but uses nothing from it.
#ifndef SBC_ORACLE_SHIM_GSL_HISTOGRAM_H
#define SBC_ORACLE_SHIM_GSL_HISTOGRAM_H
#endif
