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

// Optional fused epilogue (no reference counterpart): every kernel that writes C can apply
//   C[r, f] = act(row_scale[r] * acc[r, f] + bias[f]),  act = ReLU or identity,
// which covers the GCN layer  relu(D^-1/2 A D^-1/2 X + b)  once X has been pre-scaled by D^-1/2 on its rows
// (the column scaling of A).  All pointers are optional; default = plain SpMM.
struct Epilogue {
  const float *row_scale = nullptr;   // [num_nodes]
  const float *bias = nullptr;        // [embedding_dim]
  int32_t relu = 0;
  // model 4 (fp32 operand carried as one fp16 term): the operand was pre-multiplied by 2^s, s = 14 - floor(log2(max |x|)),
  // so that its largest value sits just below fp16's maximum; the accumulator is multiplied back by 2^-s here (exact).
  // Points at the bit pattern of max |x| written by the range pass; null everywhere else.
  const int32_t *pow2_max_bits = nullptr;
  __host__ __device__ bool any() const {
    return row_scale != nullptr || bias != nullptr || relu != 0 || pow2_max_bits != nullptr;
  }
  // 2^s of the fp16 carrier from the bit pattern of max |x| (0 -> no scaling; |s| is capped by the range pass)
  __device__ static __forceinline__ int carrier_shift(int32_t max_bits) {
    return max_bits == 0 ? 0 : 14 - (((max_bits >> 23) & 0xff) - 127);
  }
  __device__ __forceinline__ float pre_scale() const {
    if (pow2_max_bits == nullptr) return 1.f;
    return __int_as_float((127 - carrier_shift(__ldg(pow2_max_bits))) << 23);   // 2^-s
  }
  __device__ __forceinline__ float scale_of(int64_t row) const { return row_scale ? __ldg(row_scale + row) : 1.f; }
  __device__ __forceinline__ float bias_of(int32_t f) const { return bias ? __ldg(bias + f) : 0.f; }
  __device__ __forceinline__ float apply(float acc, float s, float b) const {
    const float y = fmaf(acc, s, b);
    return relu ? fmaxf(y, 0.f) : y;
  }
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
