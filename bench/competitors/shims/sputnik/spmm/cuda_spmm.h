// Stand-in for <sputnik/spmm/cuda_spmm.h>, see ../sputnik.h.
#ifndef VOLTRIX_BENCH_SPUTNIK_SPMM_SHIM_H_
#define VOLTRIX_BENCH_SPUTNIK_SPMM_SHIM_H_
#include "sputnik/sputnik.h"
#endif
