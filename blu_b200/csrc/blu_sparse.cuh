/* blu_sparse.cuh -- sparse solves and the Forrest-Tomlin update (filled in below) */
#ifndef BLU_SPARSE_CUH
#define BLU_SPARSE_CUH
#include "blu_dev_common.cuh"
#endif
