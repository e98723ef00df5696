// bmat_kernels.cuh -- GPU preprocessing: CSR -> 16-row windows with column
// compaction, 16x8 TC blocks, column index lists and mma-fragment-ordered
// bitmaps.  Output is bit-exact with the reference format
// (reference: voltrix/include/voltrix/bmat_kernels.cuh)
// but the algorithm is not the reference's: where the reference sorts every
// window on ONE host thread with a std::map (bmat_kernels.cuh:264-320) and then
// rescans every window edge once per TC block on the GPU (:21-111), this file
//
//   1. encodes every edge as a 64-bit key  [window | column | row-in-window]
//      (one thread per edge, row found by binary search in indptr),
//   2. radix-sorts the keys on the GPU (CUB, only the significant bits),
//   3. turns "first occurrence of (window, column)" head flags into compacted
//      column ranks with one prefix sum,
//   4. derives block_partition / pointer1 per window, and
//   5. scatters every edge straight into its bitmap bit (atomicOr -- commutative,
//      hence deterministic) and its hind slot.  The 512 B-per-block fp32 `hspa`
//      intermediate of the reference (7.3 GB on the Reddit-shaped config) is
//      never materialised on this path.
//
// The kernel-level entry points of the reference (preprocess / hmat_cuda /
// hmat_packed_swizzle_cuda) are kept with the same names, argument meaning and
// outputs (including the fp32 hspa tiles) for callers that drive them by hand
// (reference: tests/test_spmm_kernel.py:58-127).
#ifndef VOLTRIX_B200_BMAT_KERNELS_CUH_
#define VOLTRIX_B200_BMAT_KERNELS_CUH_

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "voltrix/common.cuh"

namespace voltrix {

// ------------------------------------------------------------------------------------------
// key layout helpers
// ------------------------------------------------------------------------------------------
struct KeyBits {
  int col_bits;   // bits holding the column id
  int win_bits;   // bits holding the window id
  __host__ __device__ int col_shift() const { return 4; }
  __host__ __device__ int win_shift() const { return 4 + col_bits; }
  __host__ __device__ int end_bit() const { return 4 + col_bits + win_bits; }
};

inline int bits_for(int64_t max_value) {  // number of bits needed to store values in [0, max_value]
  int b = 1;
  while (b < 63 && (int64_t(1) << b) <= max_value) ++b;
  return b;
}

inline KeyBits make_key_bits(int32_t num_nodes, int32_t num_cols) {
  KeyBits kb;
  kb.col_bits = num_cols > 0 ? bits_for(int64_t(num_cols) - 1) : 31;
  kb.win_bits = bits_for(ceil_div<int64_t>(num_nodes, BLK_H));
  return kb;
}

struct HeadFlagOp {
  const uint64_t *keys;
  __host__ __device__ int operator()(int64_t i) const {
    return (i == 0 || (keys[i] >> 4) != (keys[i - 1] >> 4)) ? 1 : 0;
  }
};

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
struct PreprocessWorkspace {
  uint64_t *keys_in;    // [nnz]  unsorted keys; reused as compact unique-column list after the sort
  uint64_t *keys;       // [nnz]  sorted keys
  int32_t *uidx1;       // [nnz]  1-based running count of distinct (window, column) pairs
  int32_t *ustart;      // [W]    0-based index of the window's first distinct pair
  int32_t *ucount;      // [W]    number of distinct columns of the window
  void *cub_temp;
  size_t cub_temp_bytes;
};

inline size_t preprocess_cub_temp_bytes(int64_t nnz, int32_t num_windows) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const uint64_t *)nullptr, (uint64_t *)nullptr, nnz, 0, 64);
  auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), HeadFlagOp{nullptr});
  cub::DeviceScan::InclusiveSum(nullptr, b, it, (int32_t *)nullptr, nnz);
  cub::DeviceScan::InclusiveSum(nullptr, c, (const int32_t *)nullptr, (int32_t *)nullptr, num_windows);
  size_t m = a > b ? a : b;
  return align256(m > c ? m : c);
}

inline size_t preprocess_workspace_bytes(int64_t nnz, int32_t num_nodes) {
  int32_t W = ceil_div<int32_t>(num_nodes, BLK_H);
  int64_t n = nnz > 0 ? nnz : 1;
  return align256(n * 8) * 2 + align256(n * 4) + align256(size_t(W) * 4) * 2 +
         preprocess_cub_temp_bytes(n, W) + 256;
}

inline int carve_workspace(void *ws, size_t ws_bytes, int64_t nnz, int32_t num_nodes, PreprocessWorkspace &out) {
  if (ws == nullptr || ws_bytes < preprocess_workspace_bytes(nnz, num_nodes)) return VX_ERR_WORKSPACE;
  int32_t W = ceil_div<int32_t>(num_nodes, BLK_H);
  int64_t n = nnz > 0 ? nnz : 1;
  char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  out.keys_in = reinterpret_cast<uint64_t *>(p); p += align256(n * 8);
  out.keys = reinterpret_cast<uint64_t *>(p);    p += align256(n * 8);
  out.uidx1 = reinterpret_cast<int32_t *>(p);    p += align256(n * 4);
  out.ustart = reinterpret_cast<int32_t *>(p);   p += align256(size_t(W) * 4);
  out.ucount = reinterpret_cast<int32_t *>(p);   p += align256(size_t(W) * 4);
  out.cub_temp = p;
  out.cub_temp_bytes = preprocess_cub_temp_bytes(n, W);
  return VX_OK;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

// row owning edge e: largest r with indptr[r] <= e (skips empty rows correctly).
__device__ __forceinline__ int32_t vx_row_of_edge(const int32_t *__restrict__ indptr, int32_t num_nodes, int64_t e) {
  int32_t lo = 0, hi = num_nodes;  // answer in [lo, hi)
  while (hi - lo > 1) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (int64_t(__ldg(indptr + mid)) <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// Step 1.  key = [window | column | row % 16]; also the edge_to_row output of the kernel-level API.
__global__ void vx_edge_keys_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                    int32_t num_nodes, int64_t nnz, KeyBits kb, uint64_t *__restrict__ keys,
                                    int32_t *__restrict__ edge_to_row /* nullable */) {
  int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int32_t row = vx_row_of_edge(indptr, num_nodes, e);
  uint32_t col = uint32_t(__ldg(indices + e));
  keys[e] = (uint64_t(row >> 4) << kb.win_shift()) | (uint64_t(col) << 4) | uint64_t(row & 15);
  if (edge_to_row) edge_to_row[e] = row;
}

// Step 4.  One thread per window: distinct-column count -> TC blocks.  An edgeless window still
// owns one all-zero block, exactly as the reference (bmat_kernels.cuh:250-252, 298-299).
__global__ void vx_window_blocks_kernel(const int32_t *__restrict__ indptr, int32_t num_nodes, int32_t num_windows,
                                        const int32_t *__restrict__ uidx1, int32_t *__restrict__ ustart,
                                        int32_t *__restrict__ ucount, int32_t *__restrict__ block_partition) {
  int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= num_windows) return;
  int32_t r0 = w * BLK_H;
  int32_t r1 = min(r0 + BLK_H, num_nodes);
  int32_t lo = indptr[r0], hi = indptr[r1];
  if (hi == lo) {
    ustart[w] = 0;
    ucount[w] = 0;
    block_partition[w] = 1;
  } else {
    int32_t first = uidx1[lo] - 1, last = uidx1[hi - 1] - 1;
    ustart[w] = first;
    ucount[w] = last - first + 1;
    block_partition[w] = ceil_div(last - first + 1, BLK_W);
  }
}

// Step 5.  One thread per sorted edge: set its bitmap bit, first occurrence writes the hind slot.
// Bit position follows the mma.m16n8k8 A-fragment order the reference packs into
// (bmat_kernels.cuh:180-184): word idx = (r>=8) + 2*(c>=4), bit = (r%8)*4 + c%4.
__global__ void vx_scatter_tiles_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ uidx1,
                                        const int32_t *__restrict__ ustart, const int32_t *__restrict__ pointer1,
                                        int64_t nnz, KeyBits kb, int32_t *__restrict__ hind,
                                        uint32_t *__restrict__ packed) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  uint64_t key = keys[i];
  int32_t rl = int32_t(key & 15);
  int32_t col = int32_t((key >> 4) & ((uint64_t(1) << kb.col_bits) - 1));
  int32_t w = int32_t(key >> kb.win_shift());
  int32_t u1 = uidx1[i];
  bool head = (i == 0) || (uidx1[i - 1] != u1);
  int32_t rank = (u1 - 1) - ustart[w];
  int64_t b = int64_t(pointer1[w]) + (rank >> 3);
  int32_t c = rank & 7;
  if (head) hind[b * BLK_W + c] = col;
  int32_t idx = (rl >> 3) + 2 * (c >> 2);
  int32_t beta = ((rl & 7) << 2) + (c & 3);
  atomicOr(packed + b * 4 + idx, 1u << beta);
}

// number of set bits = number of DISTINCT (row, col) pairs; nnz minus this is the duplicate count
__global__ void vx_popcount_kernel(const uint32_t *__restrict__ packed, int64_t nwords,
                                   unsigned long long *__restrict__ out) {
  unsigned long long local = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nwords; i += int64_t(gridDim.x) * blockDim.x)
    local += __popc(packed[i]);
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// compact list of distinct columns per window (for the kernel-level edge_to_column output)
__global__ void vx_unique_cols_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ uidx1,
                                      int64_t nnz, KeyBits kb, int32_t *__restrict__ ucols) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  int32_t u1 = uidx1[i];
  if (i == 0 || uidx1[i - 1] != u1)
    ucols[u1 - 1] = int32_t((keys[i] >> 4) & ((uint64_t(1) << kb.col_bits) - 1));
}

// edge_to_column[e] = rank of indices[e] among its window's distinct sorted columns
// (reference: bmat_kernels.cuh:304-307), in ORIGINAL edge order.
__global__ void vx_edge_rank_kernel(const int32_t *__restrict__ indices, const int32_t *__restrict__ edge_to_row,
                                    int64_t nnz, const int32_t *__restrict__ ucols,
                                    const int32_t *__restrict__ ustart, const int32_t *__restrict__ ucount,
                                    int32_t *__restrict__ edge_to_column) {
  int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int32_t w = edge_to_row[e] >> 4;
  int32_t col = indices[e];
  const int32_t *u = ucols + ustart[w];
  int32_t lo = 0, hi = ucount[w];
  while (lo < hi) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (uint32_t(__ldg(u + mid)) < uint32_t(col)) lo = mid + 1; else hi = mid;
  }
  edge_to_column[e] = lo;
}

// zero `elems_per_block * pointer1[W]` elements (count read on the device -> no host sync)
template <typename T>
__global__ void vx_zero_blocks_kernel(const int32_t *__restrict__ pointer1, int32_t num_windows,
                                      int64_t elems_per_block, T *__restrict__ out) {
  int64_t n = int64_t(pointer1[num_windows]) * elems_per_block;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = T(0);
}

// kernel-level hmat: one thread per edge (reference: hmat_cuda_kernel, bmat_kernels.cuh:21-111, which
// instead rescans all window edges once per TC block).
__global__ void vx_hmat_edges_kernel(const int32_t *__restrict__ edge_list, const int32_t *__restrict__ edge_to_column,
                                     const int32_t *__restrict__ edge_to_row, const int32_t *__restrict__ pointer1,
                                     int64_t nnz, float *__restrict__ hspa, int32_t *__restrict__ hind) {
  int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int32_t r = edge_to_row[e];
  int32_t col = edge_to_column[e];
  int64_t b = int64_t(pointer1[r >> 4]) + (col >> 3);
  hspa[b * (BLK_H * BLK_W) + (r & 15) * BLK_W + (col & 7)] = 1.0f;
  hind[b * BLK_W + (col & 7)] = edge_list[e];
}

// kernel-level pack: one warp per TC block, one ballot per output word
// (reference: hmat_convert_uint32_swizzle_cuda_kernel, bmat_kernels.cuh:151-193; 4 active threads per CTA).
__global__ void vx_pack_swizzle_kernel(const int32_t *__restrict__ pointer1, int32_t num_windows,
                                       const float *__restrict__ hspa, uint32_t *__restrict__ packed) {
  int64_t total = pointer1[num_windows];
  int lane = threadIdx.x & 31;
  int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t b = warp; b < total; b += nwarps) {
    const float *tile = hspa + b * (BLK_H * BLK_W);
#pragma unroll
    for (int idx = 0; idx < 4; ++idx) {
      int row = (lane >> 2) + 8 * (idx & 1);
      int col = (lane & 3) + 4 * (idx >> 1);
      float v = tile[row * BLK_W + col];
      uint32_t word = __ballot_sync(0xffffffffu, fabsf(v) > 1e-5f);
      if (lane == 0) packed[b * 4 + idx] = word;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
inline int grid_for(int64_t n, int threads) { return int(ceil_div<int64_t>(n > 0 ? n : 1, threads)); }

// Steps 1-4: sorted keys + ranks stay in the workspace, block_partition / pointer1 are written.
inline int csr_window_sort(const int32_t *indptr, const int32_t *indices, int32_t num_nodes, int64_t nnz,
                           int32_t num_cols, const PreprocessWorkspace &ws, int32_t *block_partition,
                           int32_t *pointer1, int32_t *edge_to_row, cudaStream_t stream) {
  if (num_nodes < 0 || nnz < 0) return VX_ERR_INVALID_ARG;
  // int32 offsets like the reference's: TCB <= nnz + W must fit pointer1 (shard larger matrices by rows)
  if (nnz + ceil_div<int64_t>(num_nodes, BLK_H) >= (int64_t(1) << 31)) return VX_ERR_OVERFLOW;
  int32_t W = ceil_div<int32_t>(num_nodes, BLK_H);
  KeyBits kb = make_key_bits(num_nodes, num_cols);
  VX_CUDA_TRY(cudaMemsetAsync(pointer1, 0, sizeof(int32_t), stream));
  if (W == 0) return VX_OK;
  if (nnz > 0) {
    vx_edge_keys_kernel<<<grid_for(nnz, 256), 256, 0, stream>>>(indptr, indices, num_nodes, nnz, kb, ws.keys_in,
                                                                edge_to_row);
    VX_LAUNCH_CHECK();
    size_t tb = ws.cub_temp_bytes;
    VX_CUDA_TRY(cub::DeviceRadixSort::SortKeys(ws.cub_temp, tb, (const uint64_t *)ws.keys_in, ws.keys, nnz, 4,
                                               kb.end_bit(), stream));
    auto flags = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), HeadFlagOp{ws.keys});
    tb = ws.cub_temp_bytes;
    VX_CUDA_TRY(cub::DeviceScan::InclusiveSum(ws.cub_temp, tb, flags, ws.uidx1, nnz, stream));
  }
  vx_window_blocks_kernel<<<grid_for(W, 256), 256, 0, stream>>>(indptr, num_nodes, W, ws.uidx1, ws.ustart, ws.ucount,
                                                                block_partition);
  VX_LAUNCH_CHECK();
  size_t tb = ws.cub_temp_bytes;
  VX_CUDA_TRY(cub::DeviceScan::InclusiveSum(ws.cub_temp, tb, (const int32_t *)block_partition, pointer1 + 1, W, stream));
  return VX_OK;
}

// Step 5: fill hind / hspa_packed (sized 8 / 4 entries per TC block of pointer1[W]).
inline int csr_tiles_scatter(int32_t num_nodes, int64_t nnz, int32_t num_cols, const PreprocessWorkspace &ws,
                             const int32_t *pointer1, int64_t total_blocks, int32_t *hind, uint32_t *hspa_packed,
                             int64_t *unique_nnz /* nullable, device */, cudaStream_t stream) {
  KeyBits kb = make_key_bits(num_nodes, num_cols);
  if (total_blocks > 0) {
    VX_CUDA_TRY(cudaMemsetAsync(hind, 0, size_t(total_blocks) * BLK_W * sizeof(int32_t), stream));
    VX_CUDA_TRY(cudaMemsetAsync(hspa_packed, 0, size_t(total_blocks) * 4 * sizeof(uint32_t), stream));
  }
  if (nnz > 0) {
    vx_scatter_tiles_kernel<<<grid_for(nnz, 256), 256, 0, stream>>>(ws.keys, ws.uidx1, ws.ustart, pointer1, nnz, kb,
                                                                    hind, hspa_packed);
    VX_LAUNCH_CHECK();
  }
  if (unique_nnz != nullptr) {
    VX_CUDA_TRY(cudaMemsetAsync(unique_nnz, 0, sizeof(int64_t), stream));
    if (total_blocks > 0) {
      vx_popcount_kernel<<<148 * 4, 256, 0, stream>>>(hspa_packed, total_blocks * 4,
                                                       reinterpret_cast<unsigned long long *>(unique_nnz));
      VX_LAUNCH_CHECK();
    }
  }
  return VX_OK;
}

// Kernel-level API, same outputs as the reference's voltrix::preprocess (bmat_kernels.cuh:264-320)
// but every pointer is a DEVICE pointer and the work runs on `stream`.
inline int preprocess(const int32_t *edgeList, const int32_t *nodePointer, int32_t num_nodes, int64_t num_edges,
                      int blockSize_h, int blockSize_w, int32_t *blockPartition, int32_t *edgeToColumn,
                      int32_t *edgeToRow, int32_t *Pointer1, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  if (blockSize_h != BLK_H || blockSize_w != BLK_W) return VX_ERR_UNSUPPORTED;
  PreprocessWorkspace ws;
  int rc = carve_workspace(workspace, workspace_bytes, num_edges, num_nodes, ws);
  if (rc != VX_OK) return rc;
  rc = csr_window_sort(nodePointer, edgeList, num_nodes, num_edges, /*num_cols=*/0, ws, blockPartition, Pointer1,
                       edgeToRow, stream);
  if (rc != VX_OK) return rc;
  if (num_edges > 0) {
    KeyBits kb = make_key_bits(num_nodes, 0);
    int32_t *ucols = reinterpret_cast<int32_t *>(ws.keys_in);  // unsorted keys are dead after the sort
    vx_unique_cols_kernel<<<grid_for(num_edges, 256), 256, 0, stream>>>(ws.keys, ws.uidx1, num_edges, kb, ucols);
    VX_LAUNCH_CHECK();
    vx_edge_rank_kernel<<<grid_for(num_edges, 256), 256, 0, stream>>>(edgeList, edgeToRow, num_edges, ucols, ws.ustart,
                                                                      ws.ucount, edgeToColumn);
    VX_LAUNCH_CHECK();
  }
  return VX_OK;
}

// Kernel-level API (reference: hmat_cuda, bmat_kernels.cuh:195-212).  Touches only the first
// Pointer1[W] blocks of hspa / hind, like the reference; runs on `stream`.
inline int hmat_cuda(const int32_t *nodePointer, const int32_t *edgeList, const int32_t *blockPartition,
                     const int32_t *edgeToColumn, const int32_t *edgeToRow, const int32_t *Pointer1,
                     int32_t num_row_windows, int num_nodes, int64_t num_edges, float *hspa, int32_t *hind,
                     cudaStream_t stream) {
  (void)nodePointer; (void)blockPartition; (void)num_nodes;
  if (num_row_windows <= 0) return VX_OK;
  vx_zero_blocks_kernel<float><<<1184, 256, 0, stream>>>(Pointer1, num_row_windows, BLK_H * BLK_W, hspa);
  VX_LAUNCH_CHECK();
  vx_zero_blocks_kernel<int32_t><<<296, 256, 0, stream>>>(Pointer1, num_row_windows, BLK_W, hind);
  VX_LAUNCH_CHECK();
  if (num_edges > 0) {
    vx_hmat_edges_kernel<<<grid_for(num_edges, 256), 256, 0, stream>>>(edgeList, edgeToColumn, edgeToRow, Pointer1,
                                                                       num_edges, hspa, hind);
    VX_LAUNCH_CHECK();
  }
  return VX_OK;
}

// Kernel-level API (reference: hmat_packed_swizzle_cuda, bmat_kernels.cuh:228-242).
inline int hmat_packed_swizzle_cuda(int32_t num_row_windows, const int32_t *Pointer1, const float *hspa,
                                    uint32_t *hspa_packed, cudaStream_t stream) {
  if (num_row_windows <= 0) return VX_OK;
  vx_pack_swizzle_kernel<<<148 * 8, 256, 0, stream>>>(Pointer1, num_row_windows, hspa, hspa_packed);
  VX_LAUNCH_CHECK();
  return VX_OK;
}

// ------------------------------------------------------------------------------------------
// Value tiles (no reference counterpart: the reference's format is binary, bmat_kernels.cuh:100-103).
// For a CSR matrix WITH values the tensor-core kernel reads, per TC block, a 16 x 8 tile of 16-bit values laid out
// as the shared-memory image of its A^T operand chunk: element (row r, column slot c) at (r >> 3) * 64 + (r & 7) * 8 + c
// (in elements; 128 elements = 256 bytes per block).  One thread per stored entry, in CSR order: the entry's slot is
// found by binary search for its column among the window's compacted columns in `hind` (sorted ascending; the zero
// padding after the last real column of a window's last block compares as +infinity).  A (row, col) pair stored twice
// would need an atomic add on a 16-bit value: the caller passes coalesced input (csr_preprocess counts duplicates).
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T vx_from_float(float v);
template <> __device__ __forceinline__ __half vx_from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 vx_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void vx_value_tiles_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                      const float *__restrict__ values, int32_t num_nodes, int64_t nnz,
                                      const int32_t *__restrict__ pointer1, const int32_t *__restrict__ hind,
                                      T *__restrict__ tiles, int32_t *__restrict__ not_found) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int32_t row = vx_row_of_edge(indptr, num_nodes, e);
  const int32_t w = row >> 4, col = __ldg(indices + e);
  const int64_t b0 = pointer1[w];
  const int32_t slots = (pointer1[w + 1] - int32_t(b0)) * BLK_W;
  const int32_t *u = hind + b0 * BLK_W;
  int32_t lo = 0, hi = slots;      // first slot whose column is >= col
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    const int32_t v = __ldg(u + mid);
    const bool less = (mid > 0 && v == 0) ? false : (uint32_t(v) < uint32_t(col));   // padding zeros sort last
    if (less) lo = mid + 1; else hi = mid;
  }
  if (lo >= slots || u[lo] != col) {   // the triple does not belong to this CSR matrix
    atomicAdd(not_found, 1);
    return;
  }
  const int64_t b = b0 + (lo >> 3);
  const int32_t r = row & 15, c = lo & 7;
  tiles[b * (BLK_H * BLK_W) + (r >> 3) * 64 + (r & 7) * 8 + c] = vx_from_float<T>(__ldg(values + e));
}

// tiles: T [total_blocks * 128], zeroed here; not_found (device int32, zeroed here) counts entries whose column is not in
// the tile format (a mismatched triple) -- the caller checks it.
template <typename T>
inline int value_tiles(const int32_t *indptr, const int32_t *indices, const float *values, int32_t num_nodes,
                       int64_t nnz, const int32_t *pointer1, const int32_t *hind, int64_t total_blocks, T *tiles,
                       int32_t *not_found, cudaStream_t stream) {
  if (num_nodes < 0 || nnz < 0 || total_blocks < 0 || not_found == nullptr) return VX_ERR_INVALID_ARG;
  VX_CUDA_TRY(cudaMemsetAsync(not_found, 0, sizeof(int32_t), stream));
  if (total_blocks > 0) VX_CUDA_TRY(cudaMemsetAsync(tiles, 0, size_t(total_blocks) * BLK_H * BLK_W * sizeof(T), stream));
  if (nnz > 0) {
    vx_value_tiles_kernel<T><<<grid_for(nnz, 256), 256, 0, stream>>>(indptr, indices, values, num_nodes, nnz, pointer1,
                                                                     hind, tiles, not_found);
    VX_LAUNCH_CHECK();
  }
  return VX_OK;
}

}  // namespace voltrix

#endif  // VOLTRIX_B200_BMAT_KERNELS_CUH_
