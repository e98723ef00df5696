// schedule.cuh -- the nnz-balanced work list of the persistent SpMM kernel.
//
// The reference launches one CTA per 16-row window (spmm_kernels.cuh:2028, 2058, 2089), so a hub
// window with 20x the mean number of TC blocks serialises on one SM.  Here the
// windows are turned, once per matrix and on the GPU, into a list of work items:
//
//   * every window is classified by measured density: if gathering its nnz B rows one by one
//     (CUDA-core path) moves fewer rows than `sparse_ratio` x the 16-rows-per-K-step the
//     tensor-core path would gather, its rows go to the CUDA-core row list instead;
//   * a tensor-core window with more than `cap` TC blocks is split along K into chunks of `cap`
//     blocks (cap even, so K=16 steps never straddle chunks); split chunks write partial tiles to
//     scratch slots and a fix-up pass sums them in slot order -- deterministic, no fp32 atomics;
//   * items are sorted by descending block count (LPT) so that a persistent grid striding over
//     the list ends on the smallest items.
#ifndef VOLTRIX_B200_SCHEDULE_CUH_
#define VOLTRIX_B200_SCHEDULE_CUH_

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "voltrix/common.cuh"

namespace voltrix {

struct ScheduleCounts {   // written to the device, read back by the caller (4 ints)
  int32_t num_items;      // tensor-core work items
  int32_t num_slots;      // scratch partial tiles (16 x N fp32 each)
  int32_t num_fixups;     // split windows
  int32_t num_sparse_rows;
};

struct ScheduleWorkspace {
  int32_t *n_items;   // [W+1] items per window       -> exclusive scan
  int32_t *n_slots;   // [W+1] slots per window       -> exclusive scan
  int32_t *n_fix;     // [W+1] 1 if split             -> exclusive scan
  int32_t *n_rows;    // [W+1] CUDA-core rows         -> exclusive scan
  int32_t *sort_keys_in, *sort_keys_out;   // [max_items]
  WorkItem *items_unsorted;                // [max_items]
  void *cub_temp;
  size_t cub_temp_bytes;
};

inline size_t schedule_cub_temp_bytes(int64_t max_items, int32_t W) {
  size_t a = 0, b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (const int32_t *)nullptr, (int32_t *)nullptr, W + 1);
  cub::DeviceRadixSort::SortPairsDescending(nullptr, b, (const int32_t *)nullptr, (int32_t *)nullptr,
                                            (const WorkItem *)nullptr, (WorkItem *)nullptr, max_items);
  return align256(a > b ? a : b);
}

// upper bound on the number of items for buffer sizing: every window once + one per `cap` blocks
inline int64_t schedule_max_items(int32_t num_windows, int64_t total_blocks, int32_t cap) {
  return int64_t(num_windows) + total_blocks / (cap > 0 ? cap : 1) + 1;
}

inline size_t schedule_workspace_bytes(int32_t num_windows, int64_t max_items) {
  return align256(size_t(num_windows + 1) * 4) * 4 + align256(size_t(max_items) * 4) * 2 +
         align256(size_t(max_items) * sizeof(WorkItem)) + schedule_cub_temp_bytes(max_items, num_windows) + 256;
}

inline int carve_schedule_workspace(void *ws, size_t bytes, int32_t W, int64_t max_items, ScheduleWorkspace &o) {
  if (!ws || bytes < schedule_workspace_bytes(W, max_items)) return VX_ERR_WORKSPACE;
  char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  o.n_items = (int32_t *)p; p += align256(size_t(W + 1) * 4);
  o.n_slots = (int32_t *)p; p += align256(size_t(W + 1) * 4);
  o.n_fix = (int32_t *)p;   p += align256(size_t(W + 1) * 4);
  o.n_rows = (int32_t *)p;  p += align256(size_t(W + 1) * 4);
  o.sort_keys_in = (int32_t *)p;  p += align256(size_t(max_items) * 4);
  o.sort_keys_out = (int32_t *)p; p += align256(size_t(max_items) * 4);
  o.items_unsorted = (WorkItem *)p; p += align256(size_t(max_items) * sizeof(WorkItem));
  o.cub_temp = p;
  o.cub_temp_bytes = schedule_cub_temp_bytes(max_items, W);
  return VX_OK;
}

// pass 1: classify + count.  indptr may be null (no CSR available) -> every window is tensor-core.
__global__ void vx_sched_count_kernel(const int32_t *__restrict__ pointer1, const int32_t *__restrict__ indptr,
                                      int32_t num_nodes, int32_t W, int32_t cap, float sparse_ratio, int32_t small_blocks,
                                      int32_t *__restrict__ n_items, int32_t *__restrict__ n_slots,
                                      int32_t *__restrict__ n_fix, int32_t *__restrict__ n_rows) {
  int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > W) return;
  int32_t items = 0, slots = 0, fix = 0, rows = 0;
  if (w < W) {
    int32_t cnt = pointer1[w + 1] - pointer1[w];
    bool sparse = false;
    if (indptr != nullptr && (sparse_ratio > 0.f || small_blocks > 0)) {
      int32_t r0 = w * BLK_H, r1 = min(r0 + BLK_H, num_nodes);
      int32_t nnz = indptr[r1] - indptr[r0];
      int32_t tc_rows = 16 * ((cnt + 1) >> 1);   // B rows the tensor-core path gathers (K = 16 per step)
      // small windows (at most `small_blocks` TC blocks) pay a whole work unit -- claim, metadata, pipeline fill, TMEM
      // epilogue -- for one or two K-steps: they go to the CUDA-core rows as soon as ANY gathered slot would be padding
      sparse = float(nnz) < sparse_ratio * float(tc_rows) || (cnt <= small_blocks && nnz < tc_rows);
      if (sparse) rows = r1 - r0;
    }
    if (!sparse) {
      items = ceil_div(cnt, cap);
      if (items > 1) { slots = items; fix = 1; }
    }
  }
  n_items[w] = items; n_slots[w] = slots; n_fix[w] = fix; n_rows[w] = rows;
}

// pass 2: emit items / fix-ups / sparse rows at their scanned offsets
__global__ void vx_sched_fill_kernel(const int32_t *__restrict__ pointer1, int32_t num_nodes, int32_t W, int32_t cap,
                                     const int32_t *__restrict__ o_items, const int32_t *__restrict__ o_slots,
                                     const int32_t *__restrict__ o_fix, const int32_t *__restrict__ o_rows,
                                     WorkItem *__restrict__ items, int32_t *__restrict__ keys,
                                     FixupItem *__restrict__ fixups, int32_t *__restrict__ sparse_rows,
                                     ScheduleCounts *__restrict__ counts) {
  int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) {
    counts->num_items = o_items[W];
    counts->num_slots = o_slots[W];
    counts->num_fixups = o_fix[W];
    counts->num_sparse_rows = o_rows[W];
  }
  if (w >= W) return;
  int32_t ni = o_items[w + 1] - o_items[w];
  int32_t nr = o_rows[w + 1] - o_rows[w];
  int32_t b0 = pointer1[w], cnt = pointer1[w + 1] - b0;
  for (int32_t c = 0; c < ni; ++c) {
    WorkItem it;
    it.window = w;
    it.blk_begin = b0 + c * cap;
    it.blk_count = min(cap, cnt - c * cap);
    it.slot = ni > 1 ? o_slots[w] + c : -1;
    items[o_items[w] + c] = it;
    keys[o_items[w] + c] = it.blk_count;
  }
  if (ni > 1) {
    FixupItem f; f.window = w; f.slot_begin = o_slots[w]; f.slot_count = ni; f.pad = 0;
    fixups[o_fix[w]] = f;
  }
  for (int32_t r = 0; r < nr; ++r) sparse_rows[o_rows[w] + r] = w * BLK_H + r;
}

// First phase, on `stream`: classify, count, emit.  Outputs (device): unsorted items in the workspace,
// fixups[W], sparse_rows[num_nodes], counts.  The caller reads `counts` back, then calls sort_schedule.
inline int build_schedule(const int32_t *pointer1, const int32_t *indptr /*nullable*/, int32_t num_nodes, int32_t cap,
                          float sparse_ratio, int32_t small_blocks, int64_t max_items, FixupItem *fixups,
                          int32_t *sparse_rows, ScheduleCounts *counts, void *workspace, size_t workspace_bytes,
                          cudaStream_t stream) {
  int32_t W = ceil_div<int32_t>(num_nodes, BLK_H);
  if (cap < 2) cap = 2;
  cap &= ~1;
  ScheduleWorkspace ws;
  int rc = carve_schedule_workspace(workspace, workspace_bytes, W, max_items, ws);
  if (rc != VX_OK) return rc;
  int threads = 256, grid = ceil_div(W + 1, threads);
  vx_sched_count_kernel<<<grid, threads, 0, stream>>>(pointer1, indptr, num_nodes, W, cap, sparse_ratio, small_blocks, ws.n_items,
                                                      ws.n_slots, ws.n_fix, ws.n_rows);
  VX_LAUNCH_CHECK();
  int32_t *arrs[4] = {ws.n_items, ws.n_slots, ws.n_fix, ws.n_rows};
  for (int i = 0; i < 4; ++i) {
    size_t tb = ws.cub_temp_bytes;
    VX_CUDA_TRY(cub::DeviceScan::ExclusiveSum(ws.cub_temp, tb, arrs[i], arrs[i], W + 1, stream));
  }
  vx_sched_fill_kernel<<<grid, threads, 0, stream>>>(pointer1, num_nodes, W, cap, ws.n_items, ws.n_slots, ws.n_fix,
                                                     ws.n_rows, ws.items_unsorted, ws.sort_keys_in, fixups,
                                                     sparse_rows, counts);
  VX_LAUNCH_CHECK();
  return VX_OK;
}

// Second phase, once the caller has read `counts->num_items` back: LPT order (stable radix sort by
// descending block count, so equal-sized items keep window order and the list is deterministic).
inline int sort_schedule(int32_t num_items, int32_t num_windows, int64_t max_items, WorkItem *items, void *workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
  ScheduleWorkspace ws;
  int rc = carve_schedule_workspace(workspace, workspace_bytes, num_windows, max_items, ws);
  if (rc != VX_OK) return rc;
  if (num_items <= 0) return VX_OK;
  size_t tb = ws.cub_temp_bytes;
  VX_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(ws.cub_temp, tb, (const int32_t *)ws.sort_keys_in,
                                                        ws.sort_keys_out, (const WorkItem *)ws.items_unsorted, items,
                                                        num_items, 0, 32, stream));
  return VX_OK;
}

}  // namespace voltrix

#endif  // VOLTRIX_B200_SCHEDULE_CUH_
