/* No-op HDF5 shim so the reference sources compile/link without libhdf5.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref): output is never exercised by the oracle. */
#ifndef PINC_SHIM_HDF5_H
#define PINC_SHIM_HDF5_H
#include <stdint.h>
#include <errno.h>      /* the real hdf5.h/mpi.h pull these in; io.c relies on it */
#include <sys/types.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string.h>
typedef int64_t hid_t;          /* HDF5 >= 1.10 */
typedef unsigned long long hsize_t;
typedef int herr_t;
typedef int htri_t;
#define H5P_DEFAULT 0
#define H5T_NATIVE_DOUBLE 1
#define H5T_IEEE_F64LE 2
#define H5S_SELECT_SET 0
#define H5P_DATASET_XFER 1
#define H5P_FILE_ACCESS 2
#define H5P_DATASET_CREATE 3
#define H5FD_MPIO_COLLECTIVE 1
#define H5S_UNLIMITED ((hsize_t)-1)
#define H5F_ACC_RDWR 1
#define H5F_ACC_EXCL 2
#define H5F_ACC_TRUNC 4
#define H5F_ACC_RDONLY 0
#define H5S_ALL 0
hid_t pincShimH5();  /* unprototyped on purpose: swallows any argument list */
#define H5Sclose(...) ((herr_t)pincShimH5())
#define H5Screate_simple(...) pincShimH5()
#define H5Gcreate(...) pincShimH5()
#define H5Dclose(...) ((herr_t)pincShimH5())
#define H5Pcreate(...) pincShimH5()
#define H5Gclose(...) ((herr_t)pincShimH5())
#define H5Fclose(...) ((herr_t)pincShimH5())
#define H5Sselect_hyperslab(...) ((herr_t)pincShimH5())
#define H5Pclose(...) ((herr_t)pincShimH5())
#define H5Dwrite(...) ((herr_t)pincShimH5())
#define H5Dcreate(...) pincShimH5()
#define H5Pset_dxpl_mpio(...) ((herr_t)pincShimH5())
#define H5Fcreate(...) pincShimH5()
#define H5Fopen(...) pincShimH5()
#define H5Dopen(...) pincShimH5()
#define H5Dget_space(...) pincShimH5()
#define H5Sget_simple_extent_dims(...) ((int)pincShimH5())
#define H5Pset_fapl_mpio(...) ((herr_t)pincShimH5())
#define H5Pset_chunk(...) ((herr_t)pincShimH5())
#define H5Lexists(...) ((htri_t)pincShimH5())
#define H5Fget_name(...) ((long)pincShimH5())
#define H5Dset_extent(...) ((herr_t)pincShimH5())
#define H5Dread(...) ((herr_t)pincShimH5())
#define H5Awrite(...) ((herr_t)pincShimH5())
#define H5Aexists(...) ((htri_t)pincShimH5())
#define H5Adelete(...) ((herr_t)pincShimH5())
#define H5Acreate(...) pincShimH5()
#define H5Aclose(...) ((herr_t)pincShimH5())
#endif
