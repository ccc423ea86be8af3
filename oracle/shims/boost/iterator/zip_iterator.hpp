// Oracle shim: included by the reference, nothing used.
#ifndef SBC_ORACLE_SHIM_BOOST_ZIP_ITERATOR_HPP
#define SBC_ORACLE_SHIM_BOOST_ZIP_ITERATOR_HPP
#endif
