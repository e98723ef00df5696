"""Fused GPU preprocessing and schedule construction (no reference counterpart as wrappers).

These are the kernels ``voltrix.csr_preprocess`` drives (reference orchestration:
voltrix/spmm/spmm.py:16-89).  The reference chains preprocess (CPU) -> hmat_gen -> swizzle pack through
a 512 B/block fp32 intermediate; here one sort + one scatter produce (blk_offsets, hind, hspa_packed)
directly, bit-identical, and a third step builds the nnz-balanced work list of the SpMM kernel.
"""
import torch

from ._common import check, current_stream
from .tuner import jit_tuner

# ------------------------------------------------------------------------------- CSR -> tiles
_tiles_includes = ('"voltrix/bmat_kernels.cuh"',)
_tiles_template = """
ws_query[0] = (int64_t)voltrix::preprocess_workspace_bytes(num_edges, num_nodes);
if (op == 0) { __return_code = 0; return; }
voltrix::PreprocessWorkspace ws;
__return_code = voltrix::carve_workspace(workspace, (size_t)workspace_units * 256, num_edges, num_nodes, ws);
if (__return_code != 0) return;
if (op == 1) {
  __return_code = voltrix::csr_window_sort(indptr, indices, num_nodes, (int64_t)num_edges, num_cols, ws,
                                           block_partition, pointer1, nullptr, stream);
} else {
  __return_code = voltrix::csr_tiles_scatter(num_nodes, (int64_t)num_edges, num_cols, ws, pointer1,
                                             (int64_t)total_blocks, hind, hspa_packed, unique_nnz, stream);
}
"""
_tiles_arg_defs = (
    ("op", int),
    ("indptr", torch.int),
    ("indices", torch.int),
    ("num_nodes", int),
    ("num_edges", int),
    ("num_cols", int),
    ("block_partition", torch.int),
    ("pointer1", torch.int),
    ("total_blocks", int),
    ("hind", torch.int),
    ("hspa_packed", torch.uint32),
    ("unique_nnz", torch.int64),
    ("workspace", torch.uint8),
    ("workspace_units", int),
    ("ws_query", torch.int64),
    ("stream", torch.cuda.Stream),
)


def _tiles_runtime(args):
    return jit_tuner.compile_and_tune(name="csr_tiles_kernel", keys={}, space=tuple(), includes=_tiles_includes,
                                      arg_defs=_tiles_arg_defs, template=_tiles_template, args=args)


def preprocess_workspace_bytes(num_edges: int, num_nodes: int) -> int:
    query = torch.zeros(1, dtype=torch.int64)
    args = (0, None, None, num_nodes, num_edges, 0, None, None, 0, None, None, None, None, 0, query, current_stream())
    check(_tiles_runtime(args)(*args), "csr_tiles_kernel (workspace query)")
    return int(query[0])


def csr_window_sort_kernel(indptr, indices, num_nodes: int, num_cols: int, block_partition, pointer1, workspace):
    """Phase 1: sort (window, column) keys, rank distinct columns, write block_partition / pointer1."""
    assert indptr.is_cuda and indptr.dtype == torch.int32 and indices.is_cuda and indices.dtype == torch.int32
    query = torch.zeros(1, dtype=torch.int64)
    args = (1, indptr, indices, num_nodes, indices.numel(), num_cols, block_partition, pointer1, 0, None, None, None,
            workspace, workspace.numel() // 256, query, current_stream())
    check(_tiles_runtime(args)(*args), "csr_window_sort_kernel")


def csr_tiles_scatter_kernel(num_nodes: int, num_edges: int, num_cols: int, pointer1, total_blocks: int, hind,
                             hspa_packed, unique_nnz, workspace):
    """Phase 2: zero hind / hspa_packed and scatter every edge into its bit and column slot."""
    query = torch.zeros(1, dtype=torch.int64)
    args = (2, None, None, num_nodes, num_edges, num_cols, None, pointer1, total_blocks, hind, hspa_packed, unique_nnz,
            workspace, workspace.numel() // 256, query, current_stream())
    check(_tiles_runtime(args)(*args), "csr_tiles_scatter_kernel")


# ------------------------------------------------------------------------------- schedule
_sched_includes = ('"voltrix/schedule.cuh"',)
_sched_template = """
const int32_t W = (num_nodes + BLK_H - 1) / BLK_H;
const int64_t max_items = voltrix::schedule_max_items(W, (int64_t)total_blocks, cap);
sizes[0] = max_items;
sizes[1] = (int64_t)voltrix::schedule_workspace_bytes(W, max_items);
if (op == 0) { __return_code = 0; return; }
if (op == 1) {
  __return_code = voltrix::build_schedule(pointer1, indptr, num_nodes, cap, sparse_ratio, small_blocks, max_items,
                                          reinterpret_cast<voltrix::FixupItem*>(fixups), sparse_rows,
                                          reinterpret_cast<voltrix::ScheduleCounts*>(counts), workspace,
                                          (size_t)workspace_units * 256, stream);
} else {
  __return_code = voltrix::sort_schedule(num_items, W, max_items, reinterpret_cast<voltrix::WorkItem*>(items),
                                         workspace, (size_t)workspace_units * 256, stream);
}
"""
_sched_arg_defs = (
    ("op", int),
    ("pointer1", torch.int),
    ("indptr", torch.int),
    ("num_nodes", int),
    ("total_blocks", int),
    ("cap", int),
    ("sparse_ratio", float),
    ("small_blocks", int),
    ("num_items", int),
    ("items", torch.int),
    ("fixups", torch.int),
    ("sparse_rows", torch.int),
    ("counts", torch.int),
    ("workspace", torch.uint8),
    ("workspace_units", int),
    ("sizes", torch.int64),
    ("stream", torch.cuda.Stream),
)


def _sched_runtime(args):
    return jit_tuner.compile_and_tune(name="schedule_kernel", keys={}, space=tuple(), includes=_sched_includes,
                                      arg_defs=_sched_arg_defs, template=_sched_template, args=args)


def schedule_sizes(num_nodes: int, total_blocks: int, cap: int):
    """-> (max_items, workspace_bytes)"""
    sizes = torch.zeros(2, dtype=torch.int64)
    args = (0, None, None, num_nodes, total_blocks, cap, 0.0, 0, 0, None, None, None, None, None, 0, sizes,
            current_stream())
    check(_sched_runtime(args)(*args), "schedule_kernel (size query)")
    return int(sizes[0]), int(sizes[1])


def schedule_build_kernel(pointer1, indptr, num_nodes: int, total_blocks: int, cap: int, sparse_ratio: float, fixups,
                          sparse_rows, counts, workspace, small_blocks: int = 0):
    sizes = torch.zeros(2, dtype=torch.int64)
    args = (1, pointer1, indptr, num_nodes, total_blocks, cap, float(sparse_ratio), int(small_blocks), 0, None, fixups, sparse_rows,
            counts, workspace, workspace.numel() // 256, sizes, current_stream())
    check(_sched_runtime(args)(*args), "schedule_build_kernel")


def schedule_sort_kernel(num_items: int, num_nodes: int, total_blocks: int, cap: int, items, workspace):
    sizes = torch.zeros(2, dtype=torch.int64)
    args = (2, None, None, num_nodes, total_blocks, cap, 0.0, 0, num_items, items, None, None, None, workspace,
            workspace.numel() // 256, sizes, current_stream())
    check(_sched_runtime(args)(*args), "schedule_sort_kernel")
