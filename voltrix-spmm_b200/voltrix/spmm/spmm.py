"""Public op API: ``csr_preprocess`` and ``spmm``.

Drop-in for the reference's voltrix/spmm/spmm.py:16-114 -- same names, argument order and return
values: ``csr_preprocess(indptr, indices, num_nodes) -> (blk_offsets, hspa_packed, hind)`` (three CUDA
tensors, bit-identical to the reference's) and ``spmm(blk_offsets, hspa_packed, hind, num_nodes,
num_edges, feat) -> fp32 [num_nodes, N]``.

What differs underneath:
* preprocessing runs on the GPU end to end (the reference's column compaction is a single host
  thread) and never materialises the fp32 ``hspa`` tiles;
* ``csr_preprocess`` also builds the nnz-balanced work list and keeps the CSR arrays on the device; this
  state rides along as ``hspa_packed._vx_plan`` (callers treat the triple as opaque except for setting
  ``hspa_packed.hash_tag``, reference tests/test_spmm.py:55);
* inputs may already live on the GPU; ``feat`` may be fp32, fp16 or bf16;
* every output row is written, including the ``num_nodes % 16`` tail rows the reference leaves
  uninitialised.
There is no CPU fallback: without a CUDA device these functions raise.
"""
import math
from typing import Optional

import torch

from ..jit_kernels import (
    value_tiles_kernel,
    csr_tiles_scatter_kernel,
    csr_window_sort_kernel,
    preprocess_workspace_bytes,
    schedule_build_kernel,
    schedule_sizes,
    schedule_sort_kernel,
    spmm_csr_weighted_kernel,
    spmm_kernel,
)
from ..jit_kernels._common import alloc_workspace, require_cuda

BLK_H = 16
BLK_W = 8

# A window goes to the CUDA-core row path when gathering its nnz rows one by one moves fewer than
# SPARSE_RATIO x the rows the tensor-core path would gather for it (16 per K-step).
DEFAULT_SPARSE_RATIO = 0.5
# Windows of at most this many TC blocks go to the CUDA-core rows whenever they contain any padding (see schedule.cuh).
# The rule pays on matrices with millions of tiny windows (R-MAT: -10 %, profiles/r2k_small_blocks_sweep.txt) and costs a
# second kernel launch -- 4-8 us, up to 33 % of the whole SpMM -- on small ones (protein, DD, com-amazon:
# profiles/r2ae_routing_probe.csv, r2ag_routing_probe_more.csv), so by default it applies from SMALL_BLOCKS_MIN_TCB blocks up.
DEFAULT_SMALL_BLOCKS = 8
SMALL_BLOCKS_MIN_TCB = 4_000_000
# Longest run of TC blocks one work item accumulates in the tensor core before its partial tile is handed to the fix-up pass.
MAX_CHAIN_BLOCKS = 4096


class SpmmPlan:
    """Per-matrix device state produced by ``csr_preprocess`` and consumed by ``spmm``."""

    def __init__(self):
        self.num_nodes = 0
        self.num_edges = 0
        self.total_blocks = 0
        self.unique_nnz = 0
        self.cap = 0
        self.sparse_ratio = 0.0
        self.small_blocks = 0
        self.sparse_mean_degree = -1.0   # non-zeros per CUDA-core row (picks the warp- or the group-per-row kernel)
        self.items: Optional[torch.Tensor] = None        # int32 [num_items, 4]  (window, blk_begin, blk_count, slot)
        self.fixups: Optional[torch.Tensor] = None       # int32 [num_fixups, 4] (window, slot_begin, slot_count, 0)
        self.sparse_rows: Optional[torch.Tensor] = None  # int32 [num_sparse_rows]
        self.num_items = 0
        self.num_slots = 0
        self.num_fixups = 0
        self.num_sparse_rows = 0
        self.csr_indptr: Optional[torch.Tensor] = None
        self.csr_indices: Optional[torch.Tensor] = None
        self.block_partition: Optional[torch.Tensor] = None
        self._scratch = {}

    @property
    def has_duplicates(self) -> bool:
        return self.unique_nnz != self.num_edges

    @property
    def route_is_default(self) -> bool:
        """The routing rule csr_preprocess applies by default, or none at all (no CSR arrays to route with)."""
        return self.csr_indptr is None or (self.sparse_ratio, self.small_blocks) == default_routing(self.total_blocks)

    def signature(self) -> str:
        return f"M{self.num_nodes}_E{self.num_edges}_B{self.total_blocks}_I{self.num_items}_S{self.num_sparse_rows}"

    def scratch(self, embedding_dim: int, stream_id: int = 0) -> Optional[torch.Tensor]:
        """Partial-tile buffer of the K-split windows, one per (N, CUDA stream): launches on different streams may be in
        flight together and must not share partial tiles (the reference's spmm is stream-safe)."""
        if self.num_slots == 0:
            return None
        buf = self._scratch.get((embedding_dim, stream_id))
        if buf is None:
            buf = torch.empty(self.num_slots * BLK_H * embedding_dim, dtype=torch.float32, device=self.items.device)
            self._scratch[(embedding_dim, stream_id)] = buf
        return buf

    def launch_args(self, embedding_dim: int, stream_id: int = 0):
        """The plan part of the spmm ``launch`` argument list (see jit_kernels/spmm.py::arg_defs_for)."""
        return (self.items, self.num_items, self.fixups if self.num_fixups else None, self.num_fixups,
                self.scratch(embedding_dim, stream_id), self.csr_indptr, self.csr_indices,
                self.sparse_rows if self.num_sparse_rows else None, self.num_sparse_rows,
                float(self.sparse_mean_degree))


def default_routing(total_blocks: int):
    """``(sparse_ratio, small_blocks)`` csr_preprocess applies when the caller names neither."""
    return DEFAULT_SPARSE_RATIO, (DEFAULT_SMALL_BLOCKS if total_blocks >= SMALL_BLOCKS_MIN_TCB else 0)


def _sm_count(device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def _build_schedule(plan: "SpmmPlan", pointer1: torch.Tensor, sparse_ratio: float, small_blocks: int) -> None:
    """Phase 3 of ``csr_preprocess``: the nnz-balanced schedule (work items, fix-ups, CUDA-core row list) for a routing rule.
    The tiles do not depend on the rule, so ``reschedule`` / ``tune_routing`` call this again on a finished triple."""
    num_nodes, total_blocks, dev = plan.num_nodes, plan.total_blocks, pointer1.device
    num_row_windows = math.ceil(num_nodes / BLK_H)
    indptr = plan.csr_indptr
    plan.sparse_ratio = float(sparse_ratio) if indptr is not None else 0.0
    # phase 3: nnz-balanced schedule.  A window is split along K when it alone would exceed ~1/3 of an SM's share of
    # the TC blocks (load balance: units are claimed dynamically in LPT order, so an item that size, started first, never
    # forms the tail), and in any case beyond MAX_CHAIN_BLOCKS (accuracy: the tensor core adds each K-step
    # into its fp32 accumulator with truncation, a bias that grows linearly with the length of one accumulation chain --
    # 1.3e-3 relative on a 50 000-step chain of the R-MAT hub rows, measured against an fp64-accumulating oracle (1.3e-5
    # with 1024-step chains); chunks of <= 2048 K-steps keep it below ~6e-5, and the chunks are summed in fp32
    # round-to-nearest by the fix-up pass.  4096 blocks is above the largest window of the Reddit-shaped graph, whose
    # shards therefore need no fix-up launch).
    cap = max(64, min(MAX_CHAIN_BLOCKS, (total_blocks // (_sm_count(dev) * 3)) & ~1))
    plan.cap = cap
    max_items, sched_ws_bytes = schedule_sizes(num_nodes, total_blocks, cap)
    sched_ws = alloc_workspace(sched_ws_bytes, dev)
    fixups = torch.empty((max(num_row_windows, 1), 4), dtype=torch.int32, device=dev)
    sparse_rows = torch.empty(max(num_nodes, 1), dtype=torch.int32, device=dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    plan.small_blocks = int(small_blocks) if indptr is not None else 0
    routing = plan.sparse_ratio > 0 or plan.small_blocks > 0
    schedule_build_kernel(pointer1, plan.csr_indptr if routing else None, num_nodes, total_blocks, cap,
                          plan.sparse_ratio, fixups, sparse_rows, counts, sched_ws, small_blocks=plan.small_blocks)
    plan.num_items, plan.num_slots, plan.num_fixups, plan.num_sparse_rows = (int(v) for v in counts.tolist())
    items = torch.empty((max(plan.num_items, 1), 4), dtype=torch.int32, device=dev)
    schedule_sort_kernel(plan.num_items, num_nodes, total_blocks, cap, items, sched_ws)
    plan.items = items
    plan.fixups = fixups[: max(plan.num_fixups, 1)].clone()
    plan.sparse_rows = sparse_rows[: max(plan.num_sparse_rows, 1)].clone()
    plan.sparse_mean_degree = -1.0
    if plan.num_sparse_rows > 0:
        rows = plan.sparse_rows[: plan.num_sparse_rows].long()
        plan.sparse_mean_degree = float((indptr[rows + 1] - indptr[rows]).sum().item()) / plan.num_sparse_rows
    torch.cuda.current_stream().synchronize()
    del sched_ws


def csr_preprocess(
    indptr: torch.Tensor,
    indices: torch.Tensor,
    num_nodes: int,
    sparse_ratio: float = DEFAULT_SPARSE_RATIO,
    keep_csr: bool = True,
    num_cols: Optional[int] = None,
    small_blocks: Optional[int] = None,
):
    """``num_cols`` (extension): number of columns of A when it is not square -- a row shard of a larger
    matrix has ``num_nodes`` rows but columns spanning the whole graph."""
    assert indptr.dtype == torch.int32 and indices.dtype == torch.int32
    assert indptr.numel() == num_nodes + 1
    require_cuda()
    dev = indptr.device if indptr.is_cuda else torch.device("cuda", torch.cuda.current_device())
    indptr = indptr.contiguous().to(dev, non_blocking=True)
    indices = indices.contiguous().to(dev, non_blocking=True)

    num_edges = indices.numel()
    num_row_windows = math.ceil(num_nodes / BLK_H)
    # The sort key packs [window | column | row-in-window] with just enough column bits for ``num_cols``: an id outside
    # [0, num_cols) would spill into the window field and send the scatter kernel out of bounds.  The reference's
    # std::map compaction takes any non-negative id (a rectangular A through the 3-argument signature), so the column
    # range is measured here (one reduction; the host sync below exists anyway) rather than assumed.
    if num_edges > 0:
        lo, hi = (int(v) for v in torch.aminmax(indices))
        if lo < 0:
            raise ValueError(f"csr_preprocess: negative column index {lo}")
        if num_cols is not None and hi >= int(num_cols):
            raise ValueError(f"csr_preprocess: column index {hi} out of range for num_cols={int(num_cols)}")
    else:
        hi = -1
    num_cols = int(num_cols) if num_cols is not None else max(num_nodes, hi + 1)

    # phase 1: (window, column) sort, distinct-column ranks, TC blocks per window
    workspace = alloc_workspace(preprocess_workspace_bytes(num_edges, num_nodes), dev)
    block_partition = torch.empty(num_row_windows, dtype=torch.int32, device=dev)
    pointer1 = torch.empty(num_row_windows + 1, dtype=torch.int32, device=dev)
    csr_window_sort_kernel(indptr, indices, num_nodes, num_cols, block_partition, pointer1, workspace)

    total_blocks = int(pointer1[-1].item())   # the one host sync the reference also has (spmm/spmm.py:44)

    # phase 2: bitmaps + column lists, straight from the sorted edges
    hind = torch.empty(total_blocks * BLK_W, dtype=torch.int32, device=dev)
    hspa_packed = torch.empty(total_blocks * BLK_H * BLK_W // 32, dtype=torch.uint32, device=dev)
    unique_nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    csr_tiles_scatter_kernel(num_nodes, num_edges, num_cols, pointer1, total_blocks, hind, hspa_packed, unique_nnz,
                             workspace)

    plan = SpmmPlan()
    plan.num_nodes, plan.num_edges, plan.total_blocks = num_nodes, num_edges, total_blocks
    plan.block_partition = block_partition
    plan.unique_nnz = int(unique_nnz.item())
    del workspace

    # The CUDA-core CSR path sums every stored entry, the tile format counts a duplicated (row, col)
    # once (reference bmat_kernels.cuh:102): only keep the CSR arrays when the input is coalesced.
    use_csr = keep_csr and not plan.has_duplicates
    if use_csr:
        # an edgeless matrix still needs a non-null indices pointer for the launch ABI
        plan.csr_indptr = indptr
        plan.csr_indices = indices if num_edges > 0 else torch.zeros(1, dtype=torch.int32, device=dev)
    _build_schedule(plan, pointer1, sparse_ratio if use_csr else 0.0,
                    int(default_routing(total_blocks)[1] if small_blocks is None else small_blocks) if use_csr else 0)

    hspa_packed._vx_plan = plan
    return (
        pointer1,  # blk_offsets
        hspa_packed,
        hind,
    )


# Candidate routing rules of ``tune_routing``: (sparse_ratio, small_blocks); the installed rule is always timed first.
# (0, 0) = everything on the tensor cores.
ROUTING_CANDIDATES = ((DEFAULT_SPARSE_RATIO, 0), (DEFAULT_SPARSE_RATIO, DEFAULT_SMALL_BLOCKS), (0.25, 0), (1.0, 0),
                      (DEFAULT_SPARSE_RATIO, 32), (1.0, 32), (0.0, 0))


def reschedule(blk_offsets: torch.Tensor, hspa_packed: torch.Tensor, hind: torch.Tensor,
               sparse_ratio: Optional[float] = None, small_blocks: Optional[int] = None) -> None:
    """Rebuild the work list of a preprocessed matrix under another routing rule (which windows run on the CUDA-core rows
    instead of the tensor cores; ``schedule.cuh``), in place.  The reference-format triple is untouched -- only the plan
    hung off ``hspa_packed`` changes -- so this costs the schedule kernels alone (a few ms), not a preprocessing pass.
    Without the CSR arrays (``keep_csr=False`` or duplicated entries) there is nothing to route and the call is a no-op."""
    plan = getattr(hspa_packed, "_vx_plan", None)
    if plan is None:
        raise ValueError("reschedule: this triple carries no plan (it did not come from csr_preprocess / load_preprocessed)")
    if plan.csr_indptr is None:
        return
    _build_schedule(plan, blk_offsets, plan.sparse_ratio if sparse_ratio is None else float(sparse_ratio),
                    plan.small_blocks if small_blocks is None else int(small_blocks))
    plan._scratch.clear()                         # the number of K-split slots changed
    for attr in ("_vx_fast",):                    # prepared launches bake the old work list in
        if hasattr(hspa_packed, attr):
            getattr(hspa_packed, attr).clear()


def tune_routing(blk_offsets: torch.Tensor, hspa_packed: torch.Tensor, hind: torch.Tensor, num_nodes: int, num_edges: int,
                 feat: torch.Tensor, candidates=ROUTING_CANDIDATES, iters: int = 5):
    """Pick the routing rule by measurement (the density threshold is part of the tune space).

    For every ``(sparse_ratio, small_blocks)`` candidate the plan is rescheduled, ``spmm`` runs on ``feat`` (its own kernel
    autotune included: the tune key carries the plan's shape statistics, so each rule keeps its own winner) and is timed
    with CUDA events behind a 256 MB L2 flush; the fastest rule is left installed.  Returns ``(best_rule, {rule: ms})``.
    Deterministic given the timings' order; results are unchanged (every rule computes the same sums, on different units)."""
    plan = getattr(hspa_packed, "_vx_plan", None)
    if plan is None or plan.csr_indptr is None:
        return None, {}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=feat.device)
    out = torch.empty((num_nodes, feat.shape[1]), dtype=torch.float32, device=feat.device)
    seen, timings = set(), {}
    for rule in ((plan.sparse_ratio, plan.small_blocks),) + tuple(candidates):
        rule = (float(rule[0]), int(rule[1]))
        reschedule(blk_offsets, hspa_packed, hind, *rule)
        shape = (plan.num_items, plan.num_sparse_rows, plan.num_fixups)
        if shape in seen:                         # this rule routes exactly like an earlier one
            continue
        seen.add(shape)
        spmm(blk_offsets, hspa_packed, hind, num_nodes, num_edges, feat, out=out)      # tune / load
        ts = []
        for _ in range(iters):
            flush.zero_()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            spmm(blk_offsets, hspa_packed, hind, num_nodes, num_edges, feat, out=out)
            t1.record()
            t1.synchronize()
            ts.append(t0.elapsed_time(t1))
        timings[rule] = sorted(ts)[len(ts) // 2]
    best = min(timings, key=timings.get)
    reschedule(blk_offsets, hspa_packed, hind, *best)
    return best, timings


def spmm(
    blk_offsets: torch.Tensor,  # pointer1
    hspa_packed: torch.Tensor,
    hind: torch.Tensor,
    num_nodes: int,
    num_edges: int,
    feat: torch.Tensor,
    out: Optional[torch.Tensor] = None,
    row_scale: Optional[torch.Tensor] = None,
    bias: Optional[torch.Tensor] = None,
    relu: bool = False,
    edge_weights: Optional["EdgeWeights"] = None,
):
    """Reference signature (voltrix/spmm/spmm.py:92-101) plus keyword extensions: ``out`` (caller-owned result), the
    fused epilogue ``act(row_scale[:, None] * (A @ feat) + bias)`` (see ``spmm_kernel``) and ``edge_weights``
    (``voltrix.edge_weights``): A's stored entries carry values instead of 1 -- on the tensor cores for fp16 / bf16 ``feat``."""
    num_feats = feat.shape[1]
    output = out if out is not None else torch.empty((num_nodes, num_feats), dtype=torch.float32, device=feat.device)

    spmm_kernel(
        blk_offsets,
        hspa_packed,
        hind,
        num_nodes=num_nodes,
        num_edges=num_edges,
        embedding_dim=num_feats,
        input=feat,
        output=output,
        row_scale=row_scale,
        bias=bias,
        relu=relu,
        edge_weights=edge_weights,
    )

    return output


class EdgeWeights:
    """Per-edge values of a preprocessed matrix, in the two forms the kernels read: fp32 in CSR order (CUDA-core rows) and
    16-bit value tiles -- one 16 x 8 tile per TC block, built lazily per dense-operand dtype -- for the tensor-core kernel.
    Build with ``voltrix.edge_weights``; pass to ``voltrix.spmm(..., edge_weights=w)``."""

    def __init__(self, blk_offsets, hind, indptr, indices, values, num_nodes: int, total_blocks: int):
        self._blk_offsets, self._hind, self._indptr, self._indices = blk_offsets, hind, indptr, indices
        self.csr_values = values
        self.num_nodes, self.total_blocks = num_nodes, total_blocks
        self._tiles = {}

    def tiles(self, dtype) -> torch.Tensor:
        t = self._tiles.get(dtype)
        if t is None:
            dev = self.csr_values.device
            t = torch.empty(max(self.total_blocks, 1) * BLK_H * BLK_W, dtype=dtype, device=dev)
            not_found = torch.zeros(1, dtype=torch.int32, device=dev)
            value_tiles_kernel(self._indptr, self._indices, self.csr_values, self.num_nodes, self._blk_offsets, self._hind,
                               t[: self.total_blocks * BLK_H * BLK_W], not_found)
            missing = int(not_found.item())
            if missing:
                raise ValueError(f"edge_weights: {missing} stored entries have no slot in the tile format -- the triple was "
                                 "not built from this CSR matrix")
            self._tiles[dtype] = t
        return t


def edge_weights(blk_offsets: torch.Tensor, hspa_packed: torch.Tensor, hind: torch.Tensor, indptr: torch.Tensor,
                 indices: torch.Tensor, values: torch.Tensor) -> EdgeWeights:
    """Attach a value to every stored entry of the matrix ``csr_preprocess(indptr, indices, ...)`` turned into the triple
    (no reference counterpart: its format is binary, bmat_kernels.cuh:100-103).  ``values``: ``[nnz]`` in CSR order, any
    float dtype (kept as fp32; the tensor-core path rounds them to the dense operand's fp16 / bf16 format, the CUDA-core
    rows use them as fp32).  The matrix must be coalesced (no (row, col) pair stored twice)."""
    require_cuda()
    plan = getattr(hspa_packed, "_vx_plan", None)
    if plan is None:
        raise ValueError("edge_weights needs the triple returned by voltrix.csr_preprocess (its plan rides on hspa_packed)")
    if plan.has_duplicates:
        raise ValueError("edge_weights: the matrix stores a (row, col) pair more than once; coalesce it first")
    dev = hspa_packed.device
    indptr = indptr.to(dev, non_blocking=True).contiguous()
    indices = indices.to(dev, non_blocking=True).contiguous()
    values = values.to(dev, torch.float32, non_blocking=True).contiguous()
    assert indptr.dtype == torch.int32 and indices.dtype == torch.int32
    assert indptr.numel() == plan.num_nodes + 1 and indices.numel() == plan.num_edges == values.numel()
    return EdgeWeights(blk_offsets, hind, indptr, indices, values, plan.num_nodes, plan.total_blocks)


def spmm_weighted(indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, feat: torch.Tensor,
                  out: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
                  bias: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """``act(row_scale * (A @ feat) + bias)`` for a CSR matrix WITH values (fp32 ``values[nnz]``), straight from CSR, no
    preprocessing: the vectorised CUDA-core row kernels with a multiply per gathered row.  ``feat`` fp32 / fp16 / bf16,
    fp32 accumulation and output.  Duplicate (row, col) entries add up (the binary tile path counts them once)."""
    require_cuda()
    dev = feat.device
    indptr = indptr.to(dev, non_blocking=True).contiguous()
    indices = indices.to(dev, non_blocking=True).contiguous()
    values = values.to(dev, torch.float32, non_blocking=True).contiguous()
    num_rows, num_feats = indptr.numel() - 1, feat.shape[1]
    output = out if out is not None else torch.empty((num_rows, num_feats), dtype=torch.float32, device=dev)
    spmm_csr_weighted_kernel(indptr, indices, values, num_rows, num_feats, feat, output, row_scale=row_scale, bias=bias,
                             relu=relu)
    return output


def gcn_norm(indptr: torch.Tensor) -> torch.Tensor:
    """``D^-1/2`` of a binary adjacency matrix given its CSR row pointer (isolated rows get 0), fp32, on indptr's device."""
    deg = (indptr[1:] - indptr[:-1]).to(torch.float32)
    return torch.where(deg > 0, deg.rsqrt(), torch.zeros_like(deg))


def spmm_gcn(blk_offsets, hspa_packed, hind, num_nodes: int, num_edges: int, feat: torch.Tensor, dinv: torch.Tensor,
             bias: Optional[torch.Tensor] = None, relu: bool = False, out: Optional[torch.Tensor] = None):
    """One GCN propagation ``act(D^-1/2 A D^-1/2 X + b)`` for a square binary A: the column scaling is folded into X
    (one elementwise pass over the dense operand), the row scaling, bias and activation run in the SpMM epilogue.
    The tile format has no value array (reference bmat_kernels.cuh:102): this is how normalised adjacency is served."""
    scaled = (feat.float() * dinv[:, None]).to(feat.dtype)
    return spmm(blk_offsets, hspa_packed, hind, num_nodes, num_edges, scaled, out=out, row_scale=dinv, bias=bias, relu=relu)


class HostStreamedSpMM:
    """Extension for callers whose dense operand and result live in (pinned) HOST memory.

    ``submit(feat_host, out_host)`` enqueues one SpMM: H2D copy of ``feat_host`` on a copy-in stream, the
    kernel on a compute stream, D2H copy of the result on a copy-out stream, with double-buffered device
    operands, so that the copies of step i+1 / i-1 overlap the kernel of step i (PCIe is full duplex and
    the copy engines run beside the SMs).  ``wait()`` blocks until everything submitted has landed in host
    memory.  All kernels run on ONE compute stream, in submission order: the K-split scratch of the plan is
    never shared by two launches in flight.  The reference has no counterpart (its bench keeps operands on
    the device, bench/bm_voltrix.py:15-26); every step still goes through ``spmm`` above.
    """

    def __init__(self, blk_offsets, hspa_packed, hind, num_nodes: int, num_edges: int, num_feats: int,
                 dtype=torch.float16, input_rows: Optional[int] = None, depth: int = 2,
                 shard_upload: Optional[bool] = None, group=None):
        """``shard_upload`` (multi-GPU, one process per GPU): every rank uploads only its 1/world row slice of the dense
        operand and the ranks all-gather the slices over NVLink (NCCL) on the copy-in stream, instead of each rank pulling
        the whole operand through the host links (``voltrix.distributed.upload_slice_and_all_gather``).  Default (None):
        on whenever torch.distributed is initialised with more than one rank.  Every rank must then ``submit`` the same
        number of steps (the all-gather is a collective)."""
        require_cuda()
        dev = hspa_packed.device
        self.state = (blk_offsets, hspa_packed, hind)
        self.num_nodes, self.num_edges, self.depth = num_nodes, num_edges, depth
        rows = input_rows if input_rows is not None else num_nodes
        self.rows = rows
        self.group, self.world, self.rank = group, 1, 0
        if shard_upload is None or shard_upload:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.chunk = (rows + self.world - 1) // self.world          # all-gather needs equal slices: pad the last one
        self.feat_dev = [torch.empty(self.chunk * self.world, num_feats, dtype=dtype, device=dev) for _ in range(depth)]
        self.out_dev = [torch.empty(num_nodes, num_feats, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]      # feat_dev[b] filled
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]     # kernel on buffers b finished
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]     # out_dev[b] copied out
        self.count = 0
        self.fork()

    def fork(self, stream=None):
        """Order everything submitted from now on after the work already queued on ``stream`` (default: current)."""
        stream = stream if stream is not None else torch.cuda.current_stream(self.s_in.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(stream)

    def submit(self, feat_host: torch.Tensor, out_host: torch.Tensor):
        b = self.count % self.depth
        first_use = self.count < self.depth
        with torch.cuda.stream(self.s_in):
            if not first_use:
                self.s_in.wait_event(self.ev_run[b])     # the kernel that read feat_dev[b] is done
            if self.world == 1:
                self.feat_dev[b].copy_(feat_host, non_blocking=True)
            else:
                from ..distributed import upload_slice_and_all_gather
                upload_slice_and_all_gather(self.feat_dev[b], feat_host, self.rank, self.world, self.group)
            self.ev_in[b].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[b])
            if not first_use:
                self.s_run.wait_event(self.ev_out[b])    # out_dev[b] has been copied out
            operand = self.feat_dev[b] if self.world == 1 else self.feat_dev[b][: self.rows]   # drop the all-gather padding
            spmm(*self.state, self.num_nodes, self.num_edges, operand, out=self.out_dev[b])
            self.ev_run[b].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[b])
            out_host.copy_(self.out_dev[b], non_blocking=True)
            self.ev_out[b].record(self.s_out)
        self.count += 1

    def join(self, stream=None):
        """Make ``stream`` (default: the current stream) wait for everything submitted so far."""
        stream = stream if stream is not None else torch.cuda.current_stream(self.s_in.device)
        for s in (self.s_in, self.s_run, self.s_out):
            stream.wait_stream(s)

    def wait(self):
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()


# ------------------------------------------------------------------------------------------------------------
# On-disk format of a preprocessed matrix: preprocessing + scheduling are paid once
# per graph, the autotuner's winners already persist in <cache>/tuned.json; this closes the loop across processes.
# ------------------------------------------------------------------------------------------------------------
FORMAT_VERSION = 1
_PLAN_TENSORS = ("items", "fixups", "sparse_rows", "csr_indptr", "csr_indices", "block_partition")
_PLAN_SCALARS = ("num_nodes", "num_edges", "total_blocks", "unique_nnz", "cap", "sparse_ratio", "num_items", "num_slots",
                 "num_fixups", "num_sparse_rows", "small_blocks", "sparse_mean_degree")


def save_preprocessed(path: str, blk_offsets: torch.Tensor, hspa_packed: torch.Tensor, hind: torch.Tensor,
                      hash_tag: Optional[str] = None) -> None:
    """Write the reference-format triple and the work list / CSR state hung off ``hspa_packed`` to ``path`` (torch.save of
    CPU tensors).  ``hash_tag`` (default: the one set on ``hspa_packed``, if any) is stored so that a reloaded matrix maps
    to the same autotune key."""
    plan = getattr(hspa_packed, "_vx_plan", None)
    blob = {
        "format": "voltrix-b200-tiles", "version": FORMAT_VERSION, "BLK_H": BLK_H, "BLK_W": BLK_W,
        "blk_offsets": blk_offsets.cpu(), "hspa_packed": hspa_packed.cpu().view(torch.int32), "hind": hind.cpu(),
        "hash_tag": hash_tag if hash_tag is not None else getattr(hspa_packed, "hash_tag", None),
        "plan": None,
    }
    if plan is not None:
        blob["plan"] = {"scalars": {k: getattr(plan, k) for k in _PLAN_SCALARS},
                        "tensors": {k: (getattr(plan, k).cpu() if getattr(plan, k) is not None else None)
                                    for k in _PLAN_TENSORS}}
    torch.save(blob, path)


def load_preprocessed(path: str, device=None):
    """Inverse of ``save_preprocessed``: returns ``(blk_offsets, hspa_packed, hind)`` on ``device`` (default: the current
    CUDA device) with the plan and ``hash_tag`` re-attached, ready for ``spmm``."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if blob.get("format") != "voltrix-b200-tiles" or blob.get("version") != FORMAT_VERSION:
        raise ValueError(f"{path}: not a voltrix-b200 tile file of version {FORMAT_VERSION}")
    if blob["BLK_H"] != BLK_H or blob["BLK_W"] != BLK_W:
        raise ValueError(f"{path}: tile geometry {blob['BLK_H']}x{blob['BLK_W']} does not match {BLK_H}x{BLK_W}")
    if device is None:
        require_cuda()
        device = torch.device("cuda", torch.cuda.current_device())
    blk_offsets = blob["blk_offsets"].to(device)
    hspa_packed = blob["hspa_packed"].view(torch.uint32).to(device)
    hind = blob["hind"].to(device)
    if blob["plan"] is not None:
        plan = SpmmPlan()
        for k, v in blob["plan"]["scalars"].items():
            setattr(plan, k, v)
        for k, v in blob["plan"]["tensors"].items():
            setattr(plan, k, v.to(device) if v is not None else None)
        hspa_packed._vx_plan = plan
    if blob["hash_tag"] is not None:
        hspa_packed.hash_tag = blob["hash_tag"]
    return blk_offsets, hspa_packed, hind
