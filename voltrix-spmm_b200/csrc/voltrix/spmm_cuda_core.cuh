// spmm_cuda_core.cuh -- vectorised CUDA-core SpMM paths.
//
// Two kernels, both fp32-accumulating and exact for fp32 input (no TF32 rounding):
//
//  * vx_spmm_csr_rows_kernel   -- one warp per output row, straight from CSR.  Lane groups of
//    LANES threads cover one B-row slice with 16-byte loads; 32/LANES groups walk the row's
//    non-zeros in parallel, 4 deep, and are shuffle-reduced at the end.  This is the path for
//    windows too sparse to fill an MMA tile (north_star subsystem 3): it gathers exactly
//    nnz(row) B rows, whereas a 16x8 TC block always gathers 8.
//
//  * vx_spmm_tile_rows_kernel  -- one warp per output row, from the reference tile format
//    (blk_offsets, hspa_packed, hind) only.  Used when a caller hands us nothing but the
//    reference triple (kernel-level API) and the tcgen05 path does not apply (N % 64 != 0).
//
// The reference has no CUDA-core path; its only kernels are the mma.sync pipelines at
// voltrix/include/voltrix/spmm_kernels.cuh:1458-2001.  Semantics follow SURVEY.md Appendix A:
// C[row, :] = sum over distinct columns c of row of B[c, :], fp32 accumulation.
#ifndef VOLTRIX_B200_SPMM_CUDA_CORE_CUH_
#define VOLTRIX_B200_SPMM_CUDA_CORE_CUH_

#include <type_traits>

#include "voltrix/common.cuh"

namespace voltrix {

template <typename T> struct Vec16;  // 16 bytes of T, unpacked to fp32
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void fma(float (&acc)[4], const uint4 &v, float w) {
    acc[0] = fmaf(w, __uint_as_float(v.x), acc[0]); acc[1] = fmaf(w, __uint_as_float(v.y), acc[1]);
    acc[2] = fmaf(w, __uint_as_float(v.z), acc[2]); acc[3] = fmaf(w, __uint_as_float(v.w), acc[3]);
  }
  __device__ static void add(float (&acc)[4], const uint4 &v) {
    acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y);
    acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static void fma(float (&acc)[8], const uint4 &v, float w) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      acc[2 * i] = fmaf(w, f.x, acc[2 * i]); acc[2 * i + 1] = fmaf(w, f.y, acc[2 * i + 1]);
    }
  }
  __device__ static void add(float (&acc)[8], const uint4 &v) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
    }
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void fma(float (&acc)[8], const uint4 &v, float w) {
    const uint32_t *x = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fmaf(w, __uint_as_float(x[i] << 16), acc[2 * i]);
      acc[2 * i + 1] = fmaf(w, __uint_as_float(x[i] & 0xffff0000u), acc[2 * i + 1]);
    }
  }
  __device__ static void add(float (&acc)[8], const uint4 &v) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      acc[2 * i] += __uint_as_float(w[i] << 16);
      acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

__device__ __forceinline__ uint4 vx_ldg16(const void *p) {
  return __ldg(reinterpret_cast<const uint4 *>(p));
}

// Row-slice access policies of the kernels below.  VecAccess: 16-byte loads of B, float4 stores of C (needs
// N % (16 / sizeof(T)) == 0 and 16-byte aligned B / C).  ScalarAccess: one element per lane -- any N, any alignment;
// the launchers fall back to it when the vector rule does not hold (N = 100, N = 1, odd row strides ...).
template <typename T>
struct VecAccess {
  static constexpr int N = Vec16<T>::N;
  using Reg = uint4;
  __device__ static Reg load(const T *p) { return vx_ldg16(p); }
  __device__ static void add(float (&acc)[N], const Reg &v) { Vec16<T>::add(acc, v); }
  __device__ static void fma(float (&acc)[N], const Reg &v, float w) { Vec16<T>::fma(acc, v, w); }
  __device__ static void store(float *dst, const float (&acc)[N]) {
    float4 *d = reinterpret_cast<float4 *>(dst);
#pragma unroll
    // streaming stores: C is written once and never read by this kernel -- keep L2 for the gathered B rows
    for (int i = 0; i < N / 4; ++i) __stcs(d + i, make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]));
  }
};
template <typename T>
struct ScalarAccess {
  static constexpr int N = 1;
  using Reg = float;
  __device__ static Reg load(const T *p) {
    if constexpr (sizeof(T) == 4) return __ldg(reinterpret_cast<const float *>(p));
    else if constexpr (std::is_same<T, __half>::value) return __half2float(__ldg(p));
    else return __uint_as_float(uint32_t(__ldg(reinterpret_cast<const unsigned short *>(p))) << 16);
  }
  __device__ static void add(float (&acc)[1], const Reg &v) { acc[0] += v; }
  __device__ static void fma(float (&acc)[1], const Reg &v, float w) { acc[0] = fmaf(w, v, acc[0]); }
  __device__ static void store(float *dst, const float (&acc)[1]) { dst[0] = acc[0]; }
};

// rows: either all rows [0, num_rows) (row_list == nullptr) or the rows named by
// row_list[0..num_rows).  grid.x * warps_per_block >= num_rows, grid.y = feature chunks.
// WEIGHTED: `vals[e]` multiplies the gathered row of non-zero e (general CSR values).  The
// binary instantiation is the product path of the tile format; the weighted one serves voltrix.spmm_weighted.
// RPW: rows per warp.  With RPW > 1 the warp first loads the row ids and the CSR bounds of all its rows (independent loads:
// the row_list -> indptr round trips are paid once per RPW rows) and then walks them; a launch over millions of short or
// empty rows (the sparse windows of an R-MAT matrix: 71 % of those rows have no entry) needs RPW times fewer warps.
// Measured with RPW = 1 / 2 / 4 / 8 (profiles/r2ac_csr_rows_per_warp.txt): the 21.8 M sparse rows of R-MAT-25 9.0 / 7.0 / 6.3 / 6.6 ms,
// YeastH N = 512 fp16 3.00 / 2.62 / 2.36 / 2.77 ms -- 4 everywhere.  The group-per-row kernel below already packs 2-8 rows
// into a warp; giving each group four rows made it 0-20 % slower, so it keeps one.
template <typename T, int LANES, bool WEIGHTED = false, typename A = VecAccess<T>, int RPW = 1>
__global__ void __launch_bounds__(256)
vx_spmm_csr_rows_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const int32_t *__restrict__ row_list, int32_t num_rows, int32_t N,
                   const T *__restrict__ B, float *__restrict__ C, Epilogue epi, const float *__restrict__ vals = nullptr) {
  constexpr int EPL = A::N;               // elements per lane per load
  constexpr int GROUPS = 32 / LANES;          // non-zeros processed in parallel by one warp
  constexpr int CHUNK = LANES * EPL;          // features covered by one pass
  const int lane = threadIdx.x & 31;
  const int sub = lane % LANES;               // position inside the row slice
  const int grp = lane / LANES;
  const int64_t item0 = (int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (item0 >= num_rows) return;
  const int32_t f0 = blockIdx.y * CHUNK + sub * EPL;
  const bool active = f0 < N;
  int32_t rows[RPW], begs[RPW], ends[RPW];
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
    const int64_t it = item0 + k;
    rows[k] = it < num_rows ? (row_list ? __ldg(row_list + it) : int32_t(it)) : -1;
  }
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
    begs[k] = rows[k] >= 0 ? __ldg(indptr + rows[k]) : 0;
    ends[k] = rows[k] >= 0 ? __ldg(indptr + rows[k] + 1) : 0;
  }
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
  const int32_t row = rows[k], beg = begs[k], end = ends[k];
  if (row < 0) break;                         // warp-uniform: every lane of the warp walks the same rows

  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;

  const T *Bf = B + f0;
  int32_t e = beg + grp;
  // 4 independent gathers in flight per lane group
  for (; e + 3 * GROUPS < end; e += 4 * GROUPS) {
    int32_t c0 = __ldg(indices + e), c1 = __ldg(indices + e + GROUPS);
    int32_t c2 = __ldg(indices + e + 2 * GROUPS), c3 = __ldg(indices + e + 3 * GROUPS);
    if (active) {
      const typename A::Reg v0 = A::load(Bf + int64_t(c0) * N), v1 = A::load(Bf + int64_t(c1) * N);
      const typename A::Reg v2 = A::load(Bf + int64_t(c2) * N), v3 = A::load(Bf + int64_t(c3) * N);
      if constexpr (WEIGHTED) {
        A::fma(acc, v0, __ldg(vals + e)); A::fma(acc, v1, __ldg(vals + e + GROUPS));
        A::fma(acc, v2, __ldg(vals + e + 2 * GROUPS)); A::fma(acc, v3, __ldg(vals + e + 3 * GROUPS));
      } else {
        A::add(acc, v0); A::add(acc, v1);
        A::add(acc, v2); A::add(acc, v3);
      }
    }
  }
  for (; e < end; e += GROUPS) {
    int32_t c0 = __ldg(indices + e);
    if (active) {
      if constexpr (WEIGHTED) A::fma(acc, A::load(Bf + int64_t(c0) * N), __ldg(vals + e));
      else A::add(acc, A::load(Bf + int64_t(c0) * N));
    }
  }
  // fixed-order tree reduction over the lane groups (deterministic)
#pragma unroll
  for (int off = 16; off >= LANES; off >>= 1) {
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  }
  if (grp == 0 && active) {
    if (epi.any()) {
      const float sc = epi.scale_of(row);
#pragma unroll
      for (int i = 0; i < EPL; ++i) acc[i] = epi.apply(acc[i], sc, epi.bias_of(f0 + i));
    }
    A::store(C + int64_t(row) * N + f0, acc);
  }
  }
}

// Low-degree variant: one lane GROUP (LANES threads) per row, 32 / LANES rows per warp, no cross-group reduction.
// With a mean degree of 2-10 (Yeast, DD, com-amazon ... in the C3 suite) a whole warp per row leaves most lane groups
// without a non-zero; here every group walks its own row, 4 gathers deep.
template <typename T, int LANES, bool WEIGHTED = false>
__global__ void __launch_bounds__(256)
vx_spmm_csr_subwarp_rows_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                           const int32_t *__restrict__ row_list, int32_t num_rows, int32_t N,
                           const T *__restrict__ B, float *__restrict__ C, Epilogue epi,
                           const float *__restrict__ vals = nullptr) {
  constexpr int EPL = Vec16<T>::N;
  constexpr int CHUNK = LANES * EPL;
  const int32_t item = int32_t((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / LANES);
  if (item >= num_rows) return;               // no warp-collective below: lanes may leave early
  const int sub = threadIdx.x % LANES;
  const int32_t row = row_list ? row_list[item] : item;
  const int32_t f0 = blockIdx.y * CHUNK + sub * EPL;
  if (f0 >= N) return;
  const int32_t beg = indptr[row], end = indptr[row + 1];
  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
  const T *Bf = B + f0;
  int32_t e = beg;
  for (; e + 3 < end; e += 4) {
    const int32_t c0 = __ldg(indices + e), c1 = __ldg(indices + e + 1), c2 = __ldg(indices + e + 2),
                  c3 = __ldg(indices + e + 3);
    const uint4 v0 = vx_ldg16(Bf + int64_t(c0) * N), v1 = vx_ldg16(Bf + int64_t(c1) * N);
    const uint4 v2 = vx_ldg16(Bf + int64_t(c2) * N), v3 = vx_ldg16(Bf + int64_t(c3) * N);
    if constexpr (WEIGHTED) {
      Vec16<T>::fma(acc, v0, __ldg(vals + e)); Vec16<T>::fma(acc, v1, __ldg(vals + e + 1));
      Vec16<T>::fma(acc, v2, __ldg(vals + e + 2)); Vec16<T>::fma(acc, v3, __ldg(vals + e + 3));
    } else {
      Vec16<T>::add(acc, v0); Vec16<T>::add(acc, v1);
      Vec16<T>::add(acc, v2); Vec16<T>::add(acc, v3);
    }
  }
  for (; e < end; ++e) {
    const uint4 v = vx_ldg16(Bf + int64_t(__ldg(indices + e)) * N);
    if constexpr (WEIGHTED) Vec16<T>::fma(acc, v, __ldg(vals + e));
    else Vec16<T>::add(acc, v);
  }
  if (epi.any()) {
    const float sc = epi.scale_of(row);
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = epi.apply(acc[i], sc, epi.bias_of(f0 + i));
  }
  float4 *dst = reinterpret_cast<float4 *>(C + int64_t(row) * N + f0);
#pragma unroll
  for (int i = 0; i < EPL / 4; ++i) __stcs(dst + i, make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]));
}

// One warp per row of the tile format.  Lane l scans TC block (b0 + l) of the row's window for
// its 8-bit column mask, then the warp walks the set bits together: every step all 32 lanes load
// one 512-byte slice of one B row.
template <typename T, typename A = VecAccess<T>>
__global__ void __launch_bounds__(256)
vx_spmm_tile_rows_kernel(const int32_t *__restrict__ blk_offsets, const uint32_t *__restrict__ packed,
                    const int32_t *__restrict__ hind, int32_t num_nodes, int32_t N,
                    const T *__restrict__ B, float *__restrict__ C, Epilogue epi) {
  constexpr int EPL = A::N;
  constexpr int CHUNK = 32 * EPL;
  const int lane = threadIdx.x & 31;
  const int32_t row = int32_t(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (row >= num_nodes) return;
  const int32_t w = row >> 4, r = row & 15;
  const int32_t f0 = blockIdx.y * CHUNK + lane * EPL;
  const bool active = f0 < N;
  const int64_t bb = blk_offsets[w], be = blk_offsets[w + 1];
  const int word = r >> 3, shift = (r & 7) << 2;

  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
  const T *Bf = B + f0;

  for (int64_t b0 = bb; b0 < be; b0 += 32) {
    int64_t b = b0 + lane;
    uint32_t mask = 0;
    if (b < be) {
      uint32_t lo = __ldg(packed + b * 4 + word), hi = __ldg(packed + b * 4 + word + 2);
      mask = ((lo >> shift) & 0xfu) | (((hi >> shift) & 0xfu) << 4);
    }
    uint32_t ballot = __ballot_sync(0xffffffffu, mask != 0);
    while (ballot) {
      int src = __ffs(ballot) - 1;
      ballot &= ballot - 1;
      uint32_t m = __shfl_sync(0xffffffffu, mask, src);
      const int32_t *cols = hind + (b0 + src) * BLK_W;
      while (m) {
        int c = __ffs(m) - 1;
        m &= m - 1;
        int32_t col = __ldg(cols + c);
        if (active) A::add(acc, A::load(Bf + int64_t(col) * N));
      }
    }
  }
  if (active) {
    if (epi.any()) {
      const float sc = epi.scale_of(row);
#pragma unroll
      for (int i = 0; i < EPL; ++i) acc[i] = epi.apply(acc[i], sc, epi.bias_of(f0 + i));
    }
    A::store(C + int64_t(row) * N + f0, acc);
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// 16-byte row slices need N to be a multiple of the vector width and both operands 16-byte aligned.
inline bool vec_access_ok(int32_t N, int epl, const void *B, const void *C) {
  return N % epl == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
}

constexpr float kShortRowDegree = 8.f;        // rows-per-warp launch: mean degree below this ...
constexpr int32_t kShortRowMinRows = 1 << 17;   // ... and enough rows that a quarter of the warps still fill the GPU
// mean_degree: non-zeros per row of the rows being computed (< 0 = unknown).  Rows with fewer non-zeros than a warp
// has lane groups go to the group-per-row kernel; the choice depends only on (mean_degree, N), never on timing.
template <typename T>
inline int launch_csr_rows(const int32_t *indptr, const int32_t *indices, const int32_t *row_list, int32_t num_rows,
                           int32_t N, const T *B, float *C, cudaStream_t stream, float mean_degree = -1.f,
                           const Epilogue &epi = Epilogue()) {
  constexpr int EPL = Vec16<T>::N;
  if (num_rows <= 0) return VX_OK;
  if (N <= 0) return VX_ERR_INVALID_ARG;
  dim3 block(256);
  if (!vec_access_ok(N, EPL, B, C)) {   // any N, any alignment: one element per lane
    vx_spmm_csr_rows_kernel<T, 32, false, ScalarAccess<T>><<<dim3(ceil_div(num_rows, 8), ceil_div(N, 32)), block, 0, stream>>>(
        indptr, indices, row_list, num_rows, N, B, C, epi);
    VX_LAUNCH_CHECK();
    return VX_OK;
  }
  int lanes_needed = N / EPL;  // 16-byte loads per row
  const int lanes = lanes_needed <= 4 ? 4 : lanes_needed <= 8 ? 8 : lanes_needed <= 16 ? 16 : 32;
  if (lanes < 32 && mean_degree >= 0.f && mean_degree < 4.f * float(32 / lanes)) {
    dim3 g(unsigned(ceil_div<int64_t>(int64_t(num_rows) * lanes, 256)), ceil_div(N, lanes * EPL));
    if (lanes == 4)      vx_spmm_csr_subwarp_rows_kernel<T, 4><<<g, block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
    else if (lanes == 8) vx_spmm_csr_subwarp_rows_kernel<T, 8><<<g, block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
    else                 vx_spmm_csr_subwarp_rows_kernel<T, 16><<<g, block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
    VX_LAUNCH_CHECK();
    return VX_OK;
  }
  auto grid = [&](int l) { return dim3(ceil_div(num_rows, 8), ceil_div(N, l * EPL)); };
  if (lanes == 32 && mean_degree >= 0.f && mean_degree < kShortRowDegree && num_rows >= kShortRowMinRows) {
    // millions of short rows at full row width: several rows per warp (see the kernel's RPW note)
    constexpr int RPW = 4;
    vx_spmm_csr_rows_kernel<T, 32, false, VecAccess<T>, RPW><<<dim3(ceil_div(num_rows, 8 * RPW), ceil_div(N, 32 * EPL)), block, 0,
                                                               stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
    VX_LAUNCH_CHECK();
    return VX_OK;
  }
  if (lanes == 4)       vx_spmm_csr_rows_kernel<T, 4><<<grid(4), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
  else if (lanes == 8)  vx_spmm_csr_rows_kernel<T, 8><<<grid(8), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
  else if (lanes == 16) vx_spmm_csr_rows_kernel<T, 16><<<grid(16), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
  else                  vx_spmm_csr_rows_kernel<T, 32><<<grid(32), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi);
  VX_LAUNCH_CHECK();
  return VX_OK;
}

// General CSR SpMM with fp32 values (duplicated (row, col) entries add up, as in any CSR product).  `row_list` (optional):
// compute only those rows (the sparse windows of a weighted tensor-core SpMM); num_edges < 0 = mean degree unknown.
template <typename T>
inline int launch_csr_rows_weighted(const int32_t *indptr, const int32_t *indices, const float *vals, int32_t num_rows,
                                    int64_t num_edges, int32_t N, const T *B, float *C, cudaStream_t stream,
                                    const Epilogue &epi = Epilogue(), const int32_t *row_list = nullptr) {
  constexpr int EPL = Vec16<T>::N;
  if (num_rows <= 0) return VX_OK;
  if (vals == nullptr || N <= 0) return VX_ERR_INVALID_ARG;
  if (!vec_access_ok(N, EPL, B, C)) {
    vx_spmm_csr_rows_kernel<T, 32, true, ScalarAccess<T>><<<dim3(ceil_div(num_rows, 8), ceil_div(N, 32)), dim3(256), 0, stream>>>(
        indptr, indices, row_list, num_rows, N, B, C, epi, vals);
    VX_LAUNCH_CHECK();
    return VX_OK;
  }
  const int lanes_needed = N / EPL;
  const int lanes = lanes_needed <= 4 ? 4 : lanes_needed <= 8 ? 8 : lanes_needed <= 16 ? 16 : 32;
  const float mean_degree = num_edges >= 0 ? float(num_edges) / float(num_rows) : 1e30f;
  dim3 block(256);
  if (lanes == 32 && mean_degree < kShortRowDegree && num_rows >= kShortRowMinRows) {      // RPW note above
    constexpr int RPW = 4;
    vx_spmm_csr_rows_kernel<T, 32, true, VecAccess<T>, RPW><<<dim3(ceil_div(num_rows, 8 * RPW), ceil_div(N, 32 * EPL)), block, 0,
                                                              stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
    VX_LAUNCH_CHECK();
    return VX_OK;
  }
  if (lanes < 32 && mean_degree < 4.f * float(32 / lanes)) {
    dim3 g(unsigned(ceil_div<int64_t>(int64_t(num_rows) * lanes, 256)), ceil_div(N, lanes * EPL));
    if (lanes == 4)      vx_spmm_csr_subwarp_rows_kernel<T, 4, true><<<g, block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
    else if (lanes == 8) vx_spmm_csr_subwarp_rows_kernel<T, 8, true><<<g, block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
    else                 vx_spmm_csr_subwarp_rows_kernel<T, 16, true><<<g, block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
    VX_LAUNCH_CHECK();
    return VX_OK;
  }
  auto grid = [&](int l) { return dim3(ceil_div(num_rows, 8), ceil_div(N, l * EPL)); };
  if (lanes == 4)       vx_spmm_csr_rows_kernel<T, 4, true><<<grid(4), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
  else if (lanes == 8)  vx_spmm_csr_rows_kernel<T, 8, true><<<grid(8), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
  else if (lanes == 16) vx_spmm_csr_rows_kernel<T, 16, true><<<grid(16), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
  else                  vx_spmm_csr_rows_kernel<T, 32, true><<<grid(32), block, 0, stream>>>(indptr, indices, row_list, num_rows, N, B, C, epi, vals);
  VX_LAUNCH_CHECK();
  return VX_OK;
}

template <typename T>
inline int launch_tile_rows(const int32_t *blk_offsets, const uint32_t *packed, const int32_t *hind,
                            int32_t num_nodes, int32_t N, const T *B, float *C, cudaStream_t stream,
                            const Epilogue &epi = Epilogue()) {
  constexpr int EPL = Vec16<T>::N;
  if (num_nodes <= 0) return VX_OK;
  if (N <= 0) return VX_ERR_INVALID_ARG;
  dim3 block(256), grid(ceil_div(num_nodes, 8), ceil_div(N, 32 * EPL));
  if (!vec_access_ok(N, EPL, B, C))
    vx_spmm_tile_rows_kernel<T, ScalarAccess<T>><<<dim3(grid.x, ceil_div(N, 32)), block, 0, stream>>>(
        blk_offsets, packed, hind, num_nodes, N, B, C, epi);
  else
    vx_spmm_tile_rows_kernel<T><<<grid, block, 0, stream>>>(blk_offsets, packed, hind, num_nodes, N, B, C, epi);
  VX_LAUNCH_CHECK();
  return VX_OK;
}

}  // namespace voltrix

#endif  // VOLTRIX_B200_SPMM_CUDA_CORE_CUH_
