// ref_harness.cu -- thin extern "C" shim around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  This file contains no
// algorithm of its own: it #includes the reference headers where they lie
// under $(VOLTRIX_REF)/voltrix/include (never copied into this repo) and
// forwards to the reference's own entry points:
//   voltrix::preprocess               bmat_kernels.cuh:264
//   voltrix::hmat_cuda                bmat_kernels.cuh:195
//   voltrix::hmat_packed_swizzle_cuda bmat_kernels.cuh:228
//   voltrix::voltrix_spmm_forward_cuda spmm_kernels.cuh:2003
// Built by oracle/Makefile into oracle/_ref/libvoltrix_ref.so for sm_100a
// (the reference's own flag is compute_90a, voltrix/jit/compiler.py:125; an
// sm_90a cubin does not load on a B200, the sources are untouched).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>

#include "voltrix/bmat_kernels.cuh"
#include "voltrix/spmm_kernels.cuh"

extern "C" {

// CPU. Same argument order as the reference JIT template (jit_kernels/preprocess.py:8-20).
void ref_preprocess(const int32_t *edge_list, const int32_t *node_pointer,
                    int num_nodes, int32_t *block_partition,
                    int32_t *edge_to_column, int32_t *edge_to_row,
                    int32_t *pointer1) {
  voltrix::preprocess(edge_list, node_pointer, num_nodes, BLK_H, BLK_W,
                      block_partition, edge_to_column, edge_to_row, pointer1);
}

// GPU (device pointers). Returns 0 on success.
int ref_hmat(const int32_t *node_pointer, const int32_t *edge_list,
             const int32_t *block_partition, const int32_t *edge_to_column,
             const int32_t *edge_to_row, const int32_t *pointer1,
             int num_row_windows, int num_nodes, int num_edges, float *hspa,
             int32_t *hind) {
  try {
    voltrix::hmat_cuda(node_pointer, edge_list, block_partition, edge_to_column,
                       edge_to_row, pointer1, num_row_windows, num_nodes,
                       num_edges, hspa, hind);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_hmat: %s\n", e.what());
    return 1;
  }
  return (int)cudaDeviceSynchronize();
}

int ref_hmat_packed_swizzle(int num_row_windows, const int32_t *pointer1,
                            const float *hspa, uint32_t *hspa_packed) {
  try {
    voltrix::hmat_packed_swizzle_cuda(num_row_windows, pointer1, hspa, hspa_packed);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_hmat_packed_swizzle: %s\n", e.what());
    return 1;
  }
  return (int)cudaDeviceSynchronize();
}

// Asynchronous on `stream`, like the reference (jit_kernels/spmm.py:65).
int ref_spmm(const int32_t *blk_offsets, const uint32_t *hspa_packed,
             const int32_t *hind, int num_nodes, int num_edges,
             int embedding_dim, const float *input, float *output, int model,
             void *stream) {
  try {
    voltrix::voltrix_spmm_forward_cuda(blk_offsets, hspa_packed, hind, num_nodes,
                                       num_edges, embedding_dim, input, output,
                                       model, (cudaStream_t)stream);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_spmm: %s\n", e.what());
    return 1;
  }
  return 0;
}

}  // extern "C"
