// Oracle shim (test infrastructure only): the few HDF5 C++ names the reference's sources mention
// (agents.h:25, h5tools.h:31-46, swarmdyn.cpp:553-604). The oracle never performs HDF5 I/O; the
// wrapper defines the h5tools.h functions as loud stubs.
#ifndef SBC_ORACLE_SHIM_H5CPP_H
#define SBC_ORACLE_SHIM_H5CPP_H
typedef unsigned long long hsize_t;
#define H5F_ACC_RDWR 1u
#define H5F_ACC_TRUNC 2u
namespace H5 {
class DataSet {};
class H5File {
public:
    H5File() {}
    H5File(const char *, unsigned) {}
    DataSet openDataSet(const char *) { return DataSet(); }
};
}
#endif
