/*
 * voltrix_b200.h -- C ABI of the B200-native Voltrix SpMM hot path (libvoltrix_b200.so).
 *
 * Drop-in boundary.  In the reference, the only native boundary is the JIT artefact
 * `extern "C" void launch(<args>, int& __return_code)` that voltrix/jit/template.py:104-117
 * generates around one C++ call per kernel wrapper and voltrix/jit/runtime.py:50-52 invokes
 * through ctypes.  Each entry point below is that `launch` for one wrapper, with the same
 * argument order and meaning, given a stable name, a `stream`, and an int return code
 * (0 = OK; the reference never sets its code and exit(1)s on CUDA errors,
 * spmm_kernels.cuh:39-45).  The same functions are what the JIT-generated `launch` stubs of this
 * repo call (voltrix-spmm_b200/voltrix/jit_kernels), so both routes run the same kernels.
 *
 * Conventions: every pointer is a DEVICE pointer unless named `h_*`; buffers are owned by the
 * caller; nothing is allocated or retained by the library; all work is enqueued on `stream`
 * (a cudaStream_t passed as void*, 0 = default stream) and the call returns without
 * synchronising.  int32 index arrays as in the reference; nnz < 2^31.
 *
 * Return codes: 0 OK, 1 invalid argument, 2 CUDA error, 3 workspace too small,
 * 4 unsupported configuration, 5 overflow.
 */
#ifndef VOLTRIX_B200_H_
#define VOLTRIX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_BLK_H 16 /* reference: voltrix/spmm/spmm.py:12, traits.h:6 */
#define VX_BLK_W 8  /* reference: voltrix/spmm/spmm.py:13, traits.h:7 */

/* dtype of the dense operand `input` */
#define VX_DTYPE_F32 0
#define VX_DTYPE_F16 1
#define VX_DTYPE_BF16 2

/* `model` of vx_spmm: the autotuned key (reference: jit_kernels/spmm.py:72-76, models 0/1/2) */
#define VX_MODEL_TCGEN05 0   /* tcgen05 + TMA gather4 persistent kernel (+ CUDA-core rows for sparse windows) */
#define VX_MODEL_CSR_ROWS 1  /* CUDA-core, one warp per CSR row (needs plan CSR) */
#define VX_MODEL_TILE_ROWS 2 /* CUDA-core, straight from the tile format */
#define VX_MODEL_TCGEN05_F32 3 /* fp32 input on the tcgen05 path as two bf16 terms (hi + lo); needs plan->split_ws, stages 24 */
#define VX_MODEL_TCGEN05_F32_AS_F16 4 /* fp32 input rounded to one fp16 term (11 significant bits; the reference rounds to
                                        * TF32's 10) when every value is inside fp16's normal range -- checked on the device --
                                        * else model 3's pipeline; needs plan->split_ws, plan->ticket, plan->items, stages 24 */

int vx_abi_version(void);

/* ---- a2: voltrix::preprocess (bmat_kernels.cuh:264-320; wrapper jit_kernels/preprocess.py:23) ----
 * Same outputs, computed on the GPU.  `workspace` holds vx_preprocess_workspace_bytes() bytes. */
size_t vx_preprocess_workspace_bytes(int64_t num_edges, int32_t num_nodes);
int vx_preprocess(const int32_t *edge_list, const int32_t *node_pointer, int32_t num_nodes, int64_t num_edges,
                  int32_t *block_partition, int32_t *edge_to_column, int32_t *edge_to_row, int32_t *pointer1,
                  void *workspace, size_t workspace_bytes, void *stream);

/* ---- a3: voltrix::hmat_cuda (bmat_kernels.cuh:195-212; wrapper jit_kernels/hmat_gem.py:13) ---- */
int vx_hmat_gen(const int32_t *node_pointer, const int32_t *edge_list, const int32_t *block_partition,
                const int32_t *edge_to_column, const int32_t *edge_to_row, const int32_t *pointer1,
                int32_t num_row_windows, int32_t num_nodes, int64_t num_edges, float *hspa, int32_t *hind,
                void *stream);

/* ---- a4: voltrix::hmat_packed_swizzle_cuda (bmat_kernels.cuh:228-242; wrapper jit_kernels/bmat_swizzle.py:14) */
int vx_hmat_packed_swizzle(int32_t num_row_windows, const int32_t *pointer1, const float *hspa,
                           uint32_t *hspa_packed, void *stream);

/* ---- a1: the body of csr_preprocess (voltrix/spmm/spmm.py:16-89) without the fp32 hspa detour ----
 * Phase 1 writes block_partition[W] and pointer1[W+1]; the caller reads pointer1[W] (= TC blocks),
 * allocates hind[8*TCB] and hspa_packed[4*TCB], then phase 2 fills them.  Same workspace for both.
 * num_cols: number of columns of A (0 = unknown, costs extra radix passes). */
int vx_csr_window_sort(const int32_t *indptr, const int32_t *indices, int32_t num_nodes, int64_t num_edges,
                       int32_t num_cols, int32_t *block_partition, int32_t *pointer1, void *workspace,
                       size_t workspace_bytes, void *stream);
int vx_csr_tiles_scatter(int32_t num_nodes, int64_t num_edges, int32_t num_cols, const int32_t *pointer1,
                         int64_t total_blocks, int32_t *hind, uint32_t *hspa_packed,
                         int64_t *unique_nnz /* nullable: number of distinct (row, col) pairs */, void *workspace,
                         size_t workspace_bytes, void *stream);

/* ---- nnz-balanced schedule (no reference counterpart: the reference launches one CTA per window) ---- */
typedef struct { int32_t window, blk_begin, blk_count, slot; } vx_work_item_t;        /* 16 bytes */
typedef struct { int32_t window, slot_begin, slot_count, pad; } vx_fixup_item_t;      /* 16 bytes */
typedef struct { int32_t num_items, num_slots, num_fixups, num_sparse_rows; } vx_schedule_counts_t;

int64_t vx_schedule_max_items(int32_t num_nodes, int64_t total_blocks, int32_t cap);
size_t vx_schedule_workspace_bytes(int32_t num_nodes, int64_t max_items);
/* phase 1: classify windows (indptr == NULL: all tensor-core; else a window goes to the CUDA-core rows when its nnz is
 * below sparse_ratio x the rows its TC blocks gather, or when it has at most small_blocks TC blocks and any padding at all),
 * split windows with more than `cap` blocks; fills fixups[<=W], sparse_rows[<=num_nodes], counts (device). */
int vx_schedule_build(const int32_t *pointer1, const int32_t *indptr, int32_t num_nodes, int32_t cap,
                      float sparse_ratio, int32_t small_blocks, int64_t max_items, vx_fixup_item_t *fixups, int32_t *sparse_rows,
                      vx_schedule_counts_t *counts, void *workspace, size_t workspace_bytes, void *stream);
/* phase 2 (after reading counts back): items[num_items] in LPT order. */
int vx_schedule_sort(int32_t num_items, int32_t num_nodes, int64_t max_items, vx_work_item_t *items,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ---- a5-a8: voltrix::voltrix_spmm_forward_cuda (spmm_kernels.cuh:2003-2113; wrapper jit_kernels/spmm.py:39) ----
 * Leading arguments are the reference launch's, in order: blk_offsets, hspa_packed, hind, num_nodes,
 * num_edges, embedding_dim, input, output (fp32 [num_nodes, embedding_dim], every row written --
 * including the tail rows the reference leaves uninitialised).  `plan` may be NULL (kernel-level API
 * with nothing but the reference triple). */
typedef struct {
  const vx_work_item_t *items;
  int32_t num_items;
  const vx_fixup_item_t *fixups;
  int32_t num_fixups;
  float *scratch;              /* [num_slots][16][embedding_dim] */
  const int32_t *csr_indptr;   /* coalesced CSR kept by csr_preprocess (may be NULL) */
  const int32_t *csr_indices;
  const int32_t *sparse_rows;
  int32_t num_sparse_rows;
  int64_t input_rows;          /* rows of `input` (0 = num_nodes); > num_nodes for a row shard of A */
  void *split_ws;              /* model 3 only: bf16 [input_rows][2 * embedding_dim] workspace (else NULL) */
  /* optional fused epilogue: output = act(row_scale[r] * acc + bias[f]); NULL / 0 = plain SpMM */
  const float *row_scale;      /* [num_nodes] */
  const float *bias;           /* [embedding_dim] */
  int32_t relu;
  /* ABI v5: 16 bytes of device memory.  ticket[0]: the counter the tensor-core kernel claims its work units from (atomic
   * ticket: persistent CTAs take the next unit of the LPT list when they finish one); ticket[1]: the range flag of model 4.
   * One buffer per stream that may have a launch in flight; zeroed by vx_spmm on `stream`.  NULL = static striding over
   * the list (and no model 4). */
  int32_t *ticket;
  /* ABI v5: A with a value per stored entry (NULL / NULL = the reference's binary A).  value_tiles: what vx_value_tiles
   * wrote, in `input`'s 16-bit dtype; read by model 0 in place of hspa_packed.  csr_values: fp32 [nnz] in the order of
   * csr_indices; read by model 1 and by model 0's CUDA-core rows for sparse windows. */
  const void *value_tiles;
  const float *csr_values;
  float sparse_mean_degree;    /* non-zeros per row over sparse_rows; <= 0 = unknown (warp-per-row kernel) */
} vx_plan_t;

/* `stages` (models 0, 3 and 4) = K-steps of 16 gathered rows one CTA keeps in flight; it selects a compiled variant, each with
 * its own number of producer warps and of CTAs sharing an SM: 14 (7 warps, 3 CTAs per SM; the default for any other value),
 * 22 (11, 2), 15 (5, 3), 42 (14, 1), 16 (4, 2), 12 (6, 3), 24 (8, 1).  Models 3 and 4 take 12 or 24. */
int vx_spmm(const int32_t *blk_offsets, const uint32_t *hspa_packed, const int32_t *hind, int32_t num_nodes,
            int32_t num_edges, int32_t embedding_dim, const void *input, int32_t input_dtype, float *output,
            int32_t model, int32_t stages, const vx_plan_t *plan, void *stream);

/* ---- per-edge values for the tensor-core path (no reference counterpart: the reference's tiles are binary,
 * bmat_kernels.cuh:100-103) ----
 * tiles: `tile_dtype` (VX_DTYPE_F16 / VX_DTYPE_BF16) [total_blocks * 128]: per TC block a 16 x 8 tile, element (row r,
 * column slot c) at (r / 8) * 64 + (r % 8) * 8 + c.  Zeroed, then every stored entry of the coalesced CSR matrix is
 * rounded into its slot (the slot is found by binary search in the window's hind list).  not_found (device int32[1]):
 * number of stored entries without a slot -- non-zero means the triple does not belong to this matrix. */
int vx_value_tiles(const int32_t *indptr, const int32_t *indices, const float *values, int32_t num_nodes,
                   int64_t num_edges, const int32_t *blk_offsets, const int32_t *hind, int64_t total_blocks,
                   void *tiles, int32_t tile_dtype, int32_t *not_found, void *stream);

/* ---- general CSR x dense with fp32 values (no reference counterpart: its format is binary, bmat_kernels.cuh:102) ----
 * output[r, :] = act(row_scale[r] * sum_e values[e] * input[indices[e], :] + bias), CUDA-core rows, fp32 accumulate.
 * Duplicate (row, col) entries add up.  row_scale / bias may be NULL. */
int vx_spmm_csr_weighted(const int32_t *indptr, const int32_t *indices, const float *values, int32_t num_rows,
                         int64_t num_edges, int32_t embedding_dim, const void *input, int32_t input_dtype, float *output,
                         const float *row_scale, const float *bias, int32_t relu, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VOLTRIX_B200_H_ */
