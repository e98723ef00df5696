// capi.cu -- extern "C" surface of libvoltrix_b200.so (declared in include/voltrix_b200.h).
// Thin forwarding only: every kernel lives in the header templates under csrc/voltrix/, which are
// also what the JIT-generated `launch` stubs include.
#include "voltrix_b200.h"

#include "voltrix/bmat_kernels.cuh"
#include "voltrix/schedule.cuh"
#include "voltrix/spmm_kernels.cuh"

using namespace voltrix;

static_assert(sizeof(vx_work_item_t) == sizeof(WorkItem), "ABI mismatch");
static_assert(sizeof(vx_fixup_item_t) == sizeof(FixupItem), "ABI mismatch");
static_assert(sizeof(vx_schedule_counts_t) == sizeof(ScheduleCounts), "ABI mismatch");

template <typename T>
static int spmm_dispatch(const int32_t *blk_offsets, const uint32_t *hspa_packed, const int32_t *hind,
                         int32_t num_nodes, int32_t num_edges, int32_t embedding_dim, const void *input,
                         float *output, int32_t model, int32_t stages, const SpmmPlan &plan, cudaStream_t stream) {
  const T *in = static_cast<const T *>(input);
  const bool weighted = plan.value_tiles != nullptr || plan.csr_values != nullptr;
  // `stages` = K-steps (16 gathered rows each) in flight; the number of producer warps follows from it
#define VX_TC_VARIANT(KS, NPW)                                                                                          \
  return weighted ? voltrix_spmm_forward_cuda<T, KS, NPW, true>(blk_offsets, hspa_packed, hind, num_nodes, num_edges,   \
                                                                embedding_dim, in, output, model, plan, stream)         \
                  : voltrix_spmm_forward_cuda<T, KS, NPW>(blk_offsets, hspa_packed, hind, num_nodes, num_edges,         \
                                                          embedding_dim, in, output, model, plan, stream)
  switch (stages) {
    case 12: VX_TC_VARIANT(12, 6);
    case 15: VX_TC_VARIANT(15, 5);
    case 16: VX_TC_VARIANT(16, 4);
    case 22: VX_TC_VARIANT(22, 11);
    case 24: VX_TC_VARIANT(24, 8);
    case 42: VX_TC_VARIANT(42, 14);
    default: VX_TC_VARIANT(14, 7);
  }
#undef VX_TC_VARIANT
}

template <typename T>
static int csr_weighted_dispatch(const int32_t *indptr, const int32_t *indices, const float *values, int32_t num_rows,
                                 int64_t num_edges, int32_t embedding_dim, const void *input, float *output,
                                 const Epilogue &epi, cudaStream_t stream) {
  return launch_csr_rows_weighted<T>(indptr, indices, values, num_rows, num_edges, embedding_dim,
                                     static_cast<const T *>(input), output, stream, epi);
}

extern "C" {

int vx_abi_version(void) { return 5; }

size_t vx_preprocess_workspace_bytes(int64_t num_edges, int32_t num_nodes) {
  return preprocess_workspace_bytes(num_edges, num_nodes);
}

int vx_preprocess(const int32_t *edge_list, const int32_t *node_pointer, int32_t num_nodes, int64_t num_edges,
                  int32_t *block_partition, int32_t *edge_to_column, int32_t *edge_to_row, int32_t *pointer1,
                  void *workspace, size_t workspace_bytes, void *stream) {
  return preprocess(edge_list, node_pointer, num_nodes, num_edges, BLK_H, BLK_W, block_partition, edge_to_column,
                    edge_to_row, pointer1, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vx_hmat_gen(const int32_t *node_pointer, const int32_t *edge_list, const int32_t *block_partition,
                const int32_t *edge_to_column, const int32_t *edge_to_row, const int32_t *pointer1,
                int32_t num_row_windows, int32_t num_nodes, int64_t num_edges, float *hspa, int32_t *hind,
                void *stream) {
  return hmat_cuda(node_pointer, edge_list, block_partition, edge_to_column, edge_to_row, pointer1, num_row_windows,
                   num_nodes, num_edges, hspa, hind, (cudaStream_t)stream);
}

int vx_hmat_packed_swizzle(int32_t num_row_windows, const int32_t *pointer1, const float *hspa,
                           uint32_t *hspa_packed, void *stream) {
  return hmat_packed_swizzle_cuda(num_row_windows, pointer1, hspa, hspa_packed, (cudaStream_t)stream);
}

int vx_csr_window_sort(const int32_t *indptr, const int32_t *indices, int32_t num_nodes, int64_t num_edges,
                       int32_t num_cols, int32_t *block_partition, int32_t *pointer1, void *workspace,
                       size_t workspace_bytes, void *stream) {
  PreprocessWorkspace ws;
  int rc = carve_workspace(workspace, workspace_bytes, num_edges, num_nodes, ws);
  if (rc != VX_OK) return rc;
  return csr_window_sort(indptr, indices, num_nodes, num_edges, num_cols, ws, block_partition, pointer1, nullptr,
                         (cudaStream_t)stream);
}

int vx_csr_tiles_scatter(int32_t num_nodes, int64_t num_edges, int32_t num_cols, const int32_t *pointer1,
                         int64_t total_blocks, int32_t *hind, uint32_t *hspa_packed, int64_t *unique_nnz,
                         void *workspace, size_t workspace_bytes, void *stream) {
  PreprocessWorkspace ws;
  int rc = carve_workspace(workspace, workspace_bytes, num_edges, num_nodes, ws);
  if (rc != VX_OK) return rc;
  return csr_tiles_scatter(num_nodes, num_edges, num_cols, ws, pointer1, total_blocks, hind, hspa_packed,
                           unique_nnz, (cudaStream_t)stream);
}

int64_t vx_schedule_max_items(int32_t num_nodes, int64_t total_blocks, int32_t cap) {
  return schedule_max_items(ceil_div<int32_t>(num_nodes, BLK_H), total_blocks, cap);
}

size_t vx_schedule_workspace_bytes(int32_t num_nodes, int64_t max_items) {
  return schedule_workspace_bytes(ceil_div<int32_t>(num_nodes, BLK_H), max_items);
}

int vx_schedule_build(const int32_t *pointer1, const int32_t *indptr, int32_t num_nodes, int32_t cap,
                      float sparse_ratio, int32_t small_blocks, int64_t max_items, vx_fixup_item_t *fixups,
                      int32_t *sparse_rows,
                      vx_schedule_counts_t *counts, void *workspace, size_t workspace_bytes, void *stream) {
  return build_schedule(pointer1, indptr, num_nodes, cap, sparse_ratio, small_blocks, max_items,
                        reinterpret_cast<FixupItem *>(fixups), sparse_rows,
                        reinterpret_cast<ScheduleCounts *>(counts), workspace, workspace_bytes, (cudaStream_t)stream);
}

int vx_schedule_sort(int32_t num_items, int32_t num_nodes, int64_t max_items, vx_work_item_t *items,
                     void *workspace, size_t workspace_bytes, void *stream) {
  return sort_schedule(num_items, ceil_div<int32_t>(num_nodes, BLK_H), max_items,
                       reinterpret_cast<WorkItem *>(items), workspace, workspace_bytes, (cudaStream_t)stream);
}

int vx_spmm(const int32_t *blk_offsets, const uint32_t *hspa_packed, const int32_t *hind, int32_t num_nodes,
            int32_t num_edges, int32_t embedding_dim, const void *input, int32_t input_dtype, float *output,
            int32_t model, int32_t stages, const vx_plan_t *plan, void *stream) {
  SpmmPlan p;
  if (plan) {
    p.items = reinterpret_cast<const WorkItem *>(plan->items);
    p.num_items = plan->num_items;
    p.fixups = reinterpret_cast<const FixupItem *>(plan->fixups);
    p.num_fixups = plan->num_fixups;
    p.scratch = plan->scratch;
    p.csr_indptr = plan->csr_indptr;
    p.csr_indices = plan->csr_indices;
    p.sparse_rows = plan->sparse_rows;
    p.num_sparse_rows = plan->num_sparse_rows;
    p.input_rows = plan->input_rows;
    p.split_ws = plan->split_ws;
    p.epilogue.row_scale = plan->row_scale;
    p.epilogue.bias = plan->bias;
    p.epilogue.relu = plan->relu;
    p.ticket = plan->ticket;
    p.value_tiles = plan->value_tiles;
    p.csr_values = plan->csr_values;
    p.sparse_mean_degree = plan->sparse_mean_degree > 0.f ? plan->sparse_mean_degree : -1.f;
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (input_dtype) {
    case VX_DTYPE_F32:
      return spmm_dispatch<float>(blk_offsets, hspa_packed, hind, num_nodes, num_edges, embedding_dim, input, output,
                                  model, stages, p, s);
    case VX_DTYPE_F16:
      return spmm_dispatch<__half>(blk_offsets, hspa_packed, hind, num_nodes, num_edges, embedding_dim, input, output,
                                   model, stages, p, s);
    case VX_DTYPE_BF16:
      return spmm_dispatch<__nv_bfloat16>(blk_offsets, hspa_packed, hind, num_nodes, num_edges, embedding_dim, input,
                                          output, model, stages, p, s);
  }
  return VX_ERR_INVALID_ARG;
}

int vx_value_tiles(const int32_t *indptr, const int32_t *indices, const float *values, int32_t num_nodes,
                   int64_t num_edges, const int32_t *blk_offsets, const int32_t *hind, int64_t total_blocks, void *tiles,
                   int32_t tile_dtype, int32_t *not_found, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (tile_dtype == VX_DTYPE_F16)
    return value_tiles<__half>(indptr, indices, values, num_nodes, num_edges, blk_offsets, hind, total_blocks,
                               static_cast<__half *>(tiles), not_found, s);
  if (tile_dtype == VX_DTYPE_BF16)
    return value_tiles<__nv_bfloat16>(indptr, indices, values, num_nodes, num_edges, blk_offsets, hind, total_blocks,
                                      static_cast<__nv_bfloat16 *>(tiles), not_found, s);
  return VX_ERR_INVALID_ARG;
}

int vx_spmm_csr_weighted(const int32_t *indptr, const int32_t *indices, const float *values, int32_t num_rows,
                         int64_t num_edges, int32_t embedding_dim, const void *input, int32_t input_dtype, float *output,
                         const float *row_scale, const float *bias, int32_t relu, void *stream) {
  Epilogue epi;
  epi.row_scale = row_scale;
  epi.bias = bias;
  epi.relu = relu;
  cudaStream_t s = (cudaStream_t)stream;
  switch (input_dtype) {
    case VX_DTYPE_F32: return csr_weighted_dispatch<float>(indptr, indices, values, num_rows, num_edges, embedding_dim, input, output, epi, s);
    case VX_DTYPE_F16: return csr_weighted_dispatch<__half>(indptr, indices, values, num_rows, num_edges, embedding_dim, input, output, epi, s);
    case VX_DTYPE_BF16: return csr_weighted_dispatch<__nv_bfloat16>(indptr, indices, values, num_rows, num_edges, embedding_dim, input, output, epi, s);
  }
  return VX_ERR_INVALID_ARG;
}

}  // extern "C"
