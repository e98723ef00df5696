// common.cuh -- shared definitions for the B200-native Voltrix SpMM kernels.
//
// Geometry constants keep the reference's names and values
// (reference: voltrix/include/voltrix/traits.h:6-9).
#ifndef VOLTRIX_B200_COMMON_CUH_
#define VOLTRIX_B200_COMMON_CUH_

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

#define BLK_H 16
#define BLK_W 8

namespace voltrix {

// Return codes written to `__return_code` by every JIT `launch` and returned by
// every C-ABI entry point.  The reference never sets its return code
//  and exits the process on error (spmm_kernels.cuh:39-45); here
// nothing throws or exits across the C boundary.
enum : int {
  VX_OK = 0,
  VX_ERR_INVALID_ARG = 1,
  VX_ERR_CUDA = 2,
  VX_ERR_WORKSPACE = 3,
  VX_ERR_UNSUPPORTED = 4,
  VX_ERR_OVERFLOW = 5,
};

// dtype tags of the dense operand (C ABI uses plain ints)
enum : int { VX_F32 = 0, VX_F16 = 1, VX_BF16 = 2 };

#define VX_CUDA_TRY(expr)                                                     \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      fprintf(stderr, "[voltrix] CUDA error %s at %s:%d: %s\n",               \
              cudaGetErrorName(_e), __FILE__, __LINE__, #expr);               \
      return VX_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

#define VX_LAUNCH_CHECK() VX_CUDA_TRY(cudaGetLastError())

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// One schedulable unit of the persistent SpMM kernel: a run of TC blocks of one
// 16-row window.  `slot < 0` -> the item covers the whole window and writes C
// directly; `slot >= 0` -> the window was split along K for load balance and the
// item writes a partial tile into scratch slot `slot` (summed in fixed order by
// the fix-up kernel, so results stay deterministic).
struct __align__(16) WorkItem {
  int32_t window;
  int32_t blk_begin;   // absolute TC-block index (into hind / hspa_packed)
  int32_t blk_count;
  int32_t slot;
};

// One split window for the fix-up pass.
struct __align__(16) FixupItem {
  int32_t window;
  int32_t slot_begin;
  int32_t slot_count;
  int32_t pad;
};

}  // namespace voltrix

#endif  // VOLTRIX_B200_COMMON_CUH_
