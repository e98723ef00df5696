// spmm_kernels.cuh -- the SpMM entry point behind voltrix.spmm / spmm_kernel.
//
// Mirrors the reference launcher's name and leading arguments
// (voltrix::voltrix_spmm_forward_cuda, voltrix/include/voltrix/spmm_kernels.cuh:2003-2113) so that
// the JIT template reads the same, but dispatches B200-native paths.  `model` is the autotuned key
// (the reference's three models are mma.sync tile shapes, :2014-2108); here:
//
//   model 0  tensor-core path: tcgen05 + TMA gather4 persistent kernel over the work list, plus the
//            CUDA-core row kernel for the windows the schedule classified as sparse.
//   model 1  CUDA-core CSR path for every row (needs the CSR arrays kept in the plan; exact fp32).
//   model 2  CUDA-core path straight from the tile format (needs nothing but the reference triple).
//   model 3  fp32 input on the tensor-core path: the operand is split into two bf16 terms (hi + lo, 16 mantissa
//            bits -- the reference rounds to TF32's 10) in plan.split_ws, both terms accumulate into one TMEM tile.
//   model 4  fp32 input rounded to ONE fp16 term (11 significant bits, one more than TF32) when every value lies in
//            fp16's normal range -- decided on the device by the conversion pass -- else model 3's pipeline.
//
// Unlike the reference, nothing here throws or calls exit(): errors come back as VX_* codes.
#ifndef VOLTRIX_B200_SPMM_KERNELS_CUH_
#define VOLTRIX_B200_SPMM_KERNELS_CUH_

#include <type_traits>

#include "voltrix/common.cuh"
#include "voltrix/spmm_cuda_core.cuh"
#include "voltrix/spmm_tcgen05.cuh"

namespace voltrix {

// Optional per-matrix state produced by csr_preprocess (all device pointers; any may be null).
struct SpmmPlan {
  const WorkItem *items = nullptr;       // LPT-sorted tensor-core work list
  int32_t num_items = 0;
  const FixupItem *fixups = nullptr;     // K-split windows
  int32_t num_fixups = 0;
  float *scratch = nullptr;              // [num_slots][16][N] fp32 partial tiles
  const int32_t *csr_indptr = nullptr;   // coalesced CSR (for the CUDA-core row path)
  const int32_t *csr_indices = nullptr;
  const int32_t *sparse_rows = nullptr;  // rows of the windows routed to the CUDA-core path
  int32_t num_sparse_rows = 0;
  float sparse_mean_degree = -1.f;       // non-zeros per row over sparse_rows (< 0 = unknown): picks warp- vs group-per-row
  int64_t input_rows = 0;                // rows of the dense operand (0 = num_nodes, i.e. square A);
                                         // differs for a row shard of A, whose columns span the full matrix
  void *split_ws = nullptr;              // model 3: bf16 [input_rows][2 * embedding_dim] workspace
  Epilogue epilogue;                     // optional fused row scale / bias / ReLU (default: none)
  const void *value_tiles = nullptr;     // WEIGHTED: 16-bit value tiles, 256 B per TC block (bmat_kernels.cuh::value_tiles)
  const float *csr_values = nullptr;     // WEIGHTED: fp32 values in CSR order (CUDA-core rows: sparse windows, model 1)
  int32_t *ticket = nullptr;             // 4 bytes the tensor-core kernel claims its work units from (atomic ticket; one
                                         // per stream in flight); null = static striding over the work list
};

template <typename T> struct TcSupported { static constexpr bool value = false; };
template <> struct TcSupported<__half> { static constexpr bool value = true; };
template <> struct TcSupported<__nv_bfloat16> { static constexpr bool value = true; };

// A with per-edge values (no reference counterpart): model 0 runs the WEIGHTED instantiation of
// the tensor-core kernel on plan.value_tiles (+ weighted CUDA-core rows for the sparse windows), model 1 the weighted
// CUDA-core rows on plan.csr_values.  16-bit dense operands ride the tensor cores; fp32 operands use model 1.
template <typename T, int STAGES, int NPW, int FT = 128>
inline int voltrix_spmm_weighted_forward_cuda(const int32_t *blks_offsets, const int32_t *hind, int num_nodes,
                                              int num_edges, int embedding_dim, const T *input, float *output, int model,
                                              const SpmmPlan &plan, cudaStream_t stream) {
  const int64_t b_rows = plan.input_rows > 0 ? plan.input_rows : num_nodes;
  if (model == 0) {
    if constexpr (TcSupported<T>::value) {
      if (plan.items == nullptr || plan.value_tiles == nullptr) return VX_ERR_INVALID_ARG;
      if (plan.num_fixups > 0 && plan.scratch == nullptr) return VX_ERR_INVALID_ARG;
      int rc = launch_spmm_tc<T, STAGES, NPW, 1, true, FT>(plan.items, plan.num_items, plan.fixups, plan.num_fixups,
                                                       blks_offsets, static_cast<const uint32_t *>(plan.value_tiles), hind,
                                                       num_nodes, b_rows, embedding_dim, input, output, plan.scratch,
                                                       stream, plan.epilogue, plan.ticket);
      if (rc != VX_OK) return rc;
      if (plan.num_sparse_rows > 0) {
        if (!plan.csr_indptr || !plan.csr_indices || !plan.sparse_rows || !plan.csr_values) return VX_ERR_INVALID_ARG;
        const int64_t sparse_nnz = plan.sparse_mean_degree >= 0.f
                                       ? int64_t(plan.sparse_mean_degree * float(plan.num_sparse_rows)) : int64_t(-1);
        rc = launch_csr_rows_weighted<T>(plan.csr_indptr, plan.csr_indices, plan.csr_values, plan.num_sparse_rows, sparse_nnz,
                                         embedding_dim, input, output, stream, plan.epilogue, plan.sparse_rows);
      }
      return rc;
    } else {
      return VX_ERR_UNSUPPORTED;
    }
  } else if (model == 1) {
    if (!plan.csr_indptr || !plan.csr_indices || !plan.csr_values) return VX_ERR_INVALID_ARG;
    return launch_csr_rows_weighted<T>(plan.csr_indptr, plan.csr_indices, plan.csr_values, num_nodes, num_edges,
                                       embedding_dim, input, output, stream, plan.epilogue);
  }
  return VX_ERR_UNSUPPORTED;
}

// FT (feature tile = MMA M, 128 or 64) applies to model 0; the 64-wide tile is for embedding_dim <= 64.
template <typename T, int STAGES = 32, int NPW = 8, bool WEIGHTED = false, int FT = 128>
inline int voltrix_spmm_forward_cuda(const int32_t *blks_offsets, const uint32_t *hspa_packed, const int32_t *hind,
                                     int num_nodes, int num_edges, int embedding_dim, const T *input, float *output,
                                     int model, const SpmmPlan &plan, cudaStream_t stream) {
  if (num_nodes < 0 || embedding_dim <= 0) return VX_ERR_INVALID_ARG;
  if (num_nodes == 0) return VX_OK;
  if constexpr (WEIGHTED)
    return voltrix_spmm_weighted_forward_cuda<T, STAGES, NPW, FT>(blks_offsets, hind, num_nodes, num_edges, embedding_dim,
                                                              input, output, model, plan, stream);
  const int32_t W = ceil_div<int32_t>(num_nodes, BLK_H);
  const int64_t b_rows = plan.input_rows > 0 ? plan.input_rows : num_nodes;
  if (model == 0) {
    if constexpr (TcSupported<T>::value) {
      int rc;
      if (plan.items != nullptr) {
        if (plan.num_fixups > 0 && plan.scratch == nullptr) return VX_ERR_INVALID_ARG;
        rc = launch_spmm_tc<T, STAGES, NPW, 1, false, FT>(plan.items, plan.num_items, plan.fixups, plan.num_fixups, blks_offsets,
                                       hspa_packed, hind, num_nodes, b_rows, embedding_dim, input,
                                       output, plan.scratch, stream, plan.epilogue, plan.ticket);
        if (rc != VX_OK) return rc;
        if (plan.num_sparse_rows > 0) {
          if (!plan.csr_indptr || !plan.csr_indices || !plan.sparse_rows) return VX_ERR_INVALID_ARG;
          rc = launch_csr_rows<T>(plan.csr_indptr, plan.csr_indices, plan.sparse_rows, plan.num_sparse_rows,
                                  embedding_dim, input, output, stream, plan.sparse_mean_degree, plan.epilogue);
        }
      } else {
        rc = launch_spmm_tc<T, STAGES, NPW, 1, false, FT>(nullptr, W, nullptr, 0, blks_offsets, hspa_packed, hind, num_nodes, b_rows,
                                       embedding_dim, input, output, nullptr, stream, plan.epilogue, plan.ticket);
      }
      return rc;
    } else {
      return VX_ERR_UNSUPPORTED;
    }
  } else if (model == 1) {
    if (!plan.csr_indptr || !plan.csr_indices) return VX_ERR_INVALID_ARG;
    return launch_csr_rows<T>(plan.csr_indptr, plan.csr_indices, nullptr, num_nodes, embedding_dim, input, output,
                              stream, float(num_edges) / float(num_nodes), plan.epilogue);
  } else if (model == 2) {
    return launch_tile_rows<T>(blks_offsets, hspa_packed, hind, num_nodes, embedding_dim, input, output, stream,
                               plan.epilogue);
  } else if (model == 3) {
    if constexpr (std::is_same<T, float>::value && tc_smem_bytes<STAGES, NPW, 2>() <= 227 * 1024) {
      if (plan.split_ws == nullptr) return VX_ERR_INVALID_ARG;
      if (embedding_dim % 8 != 0) return VX_ERR_UNSUPPORTED;
      if (plan.num_fixups > 0 && plan.scratch == nullptr) return VX_ERR_INVALID_ARG;
      __nv_bfloat16 *terms = static_cast<__nv_bfloat16 *>(plan.split_ws);
      int rc = launch_split_bf16x2(input, terms, b_rows, embedding_dim, stream);
      if (rc != VX_OK) return rc;
      if (plan.items != nullptr) {
        rc = launch_spmm_tc<__nv_bfloat16, STAGES, NPW, 2>(plan.items, plan.num_items, plan.fixups, plan.num_fixups,
                                                           blks_offsets, hspa_packed, hind, num_nodes, b_rows,
                                                           embedding_dim, terms, output, plan.scratch, stream,
                                                           plan.epilogue, plan.ticket);
        if (rc != VX_OK) return rc;
        if (plan.num_sparse_rows > 0) {   // sparse windows: exact fp32 rows from the original operand
          if (!plan.csr_indptr || !plan.csr_indices || !plan.sparse_rows) return VX_ERR_INVALID_ARG;
          rc = launch_csr_rows<T>(plan.csr_indptr, plan.csr_indices, plan.sparse_rows, plan.num_sparse_rows,
                                  embedding_dim, input, output, stream, plan.sparse_mean_degree, plan.epilogue);
        }
      } else {
        rc = launch_spmm_tc<__nv_bfloat16, STAGES, NPW, 2>(nullptr, W, nullptr, 0, blks_offsets, hspa_packed, hind,
                                                           num_nodes, b_rows, embedding_dim, terms, output, nullptr,
                                                           stream, plan.epilogue, plan.ticket);
      }
      return rc;
    } else {
      return VX_ERR_UNSUPPORTED;
    }
  } else if (model == 4) {
    // fp32 input at the reference's precision class, at fp16 cost: the operand is rounded to fp16 (11 significant bits; the
    // reference rounds to TF32's 10) and runs the one-term tensor-core kernel -- unless the operand does not fit fp16's
    // normal range after a power-of-two scaling (Inf / NaN, or more than a stray value in 10^6 below it), which the
    // conversion pass detects; then the gated two-term bf16 pipeline of model 3 runs instead.  Both pipelines are
    // enqueued, the device flag (plan.ticket[1]) picks one: no host round trip, CUDA-graph capturable.
    if constexpr (std::is_same<T, float>::value && tc_smem_bytes<STAGES, NPW, 2>() <= 227 * 1024) {
      if (plan.split_ws == nullptr || plan.ticket == nullptr || plan.items == nullptr) return VX_ERR_INVALID_ARG;
      if (embedding_dim % 8 != 0) return VX_ERR_UNSUPPORTED;
      if (plan.num_fixups > 0 && plan.scratch == nullptr) return VX_ERR_INVALID_ARG;
      int32_t *flag = plan.ticket + 1;       // ticket[1..3]: flag, bits of max |x|, tail-group count
      __half *as_half = static_cast<__half *>(plan.split_ws);
      __nv_bfloat16 *terms = static_cast<__nv_bfloat16 *>(plan.split_ws);
      // one memset for the whole call: the unit ticket (only one of the two gated tensor-core launches ever claims
      // from it) and the three state words of the conversion pass
      VX_CUDA_TRY(cudaMemsetAsync(plan.ticket, 0, 4 * sizeof(int32_t), stream));
      int rc = launch_cvt_f16(input, as_half, b_rows, embedding_dim, flag, stream, /*state_zeroed=*/true);
      if (rc != VX_OK) return rc;
      Epilogue carried = plan.epilogue;      // the fp16 carrier is 2^s times the operand: scale the accumulator back
      carried.pow2_max_bits = flag + 1;
      if (embedding_dim <= 64)   // the 64-wide feature tile, as for 16-bit operands of that width
        rc = launch_spmm_tc<__half, 20, 10, 1, false, 64>(plan.items, plan.num_items, plan.fixups, plan.num_fixups,
                                                          blks_offsets, hspa_packed, hind, num_nodes, b_rows, embedding_dim,
                                                          as_half, output, plan.scratch, stream, carried, plan.ticket, flag, 0,
                                                          /*reset_ticket=*/false);
      else
        rc = launch_spmm_tc<__half, 14, 7, 1>(plan.items, plan.num_items, plan.fixups, plan.num_fixups, blks_offsets,
                                              hspa_packed, hind, num_nodes, b_rows, embedding_dim, as_half, output,
                                              plan.scratch, stream, carried, plan.ticket, flag, 0, /*reset_ticket=*/false);
      if (rc != VX_OK) return rc;
      rc = launch_split_bf16x2(input, terms, b_rows, embedding_dim, stream, flag, 1);
      if (rc != VX_OK) return rc;
      rc = launch_spmm_tc<__nv_bfloat16, STAGES, NPW, 2>(plan.items, plan.num_items, plan.fixups, plan.num_fixups,
                                                         blks_offsets, hspa_packed, hind, num_nodes, b_rows, embedding_dim,
                                                         terms, output, plan.scratch, stream, plan.epilogue, plan.ticket,
                                                         flag, 1, /*reset_ticket=*/false);
      if (rc != VX_OK) return rc;
      if (plan.num_sparse_rows > 0) {   // sparse windows: exact fp32 rows from the original operand
        if (!plan.csr_indptr || !plan.csr_indices || !plan.sparse_rows) return VX_ERR_INVALID_ARG;
        rc = launch_csr_rows<T>(plan.csr_indptr, plan.csr_indices, plan.sparse_rows, plan.num_sparse_rows, embedding_dim,
                                input, output, stream, plan.sparse_mean_degree, plan.epilogue);
      }
      return rc;
    } else {
      return VX_ERR_UNSUPPORTED;
    }
  }
  return VX_ERR_INVALID_ARG;
}

}  // namespace voltrix

#endif  // VOLTRIX_B200_SPMM_KERNELS_CUH_
