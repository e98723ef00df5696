"""Row-sharded multi-GPU SpMM (new functionality: the reference is single-GPU).

Output rows are independent and a 16-row window is the atomic unit, so A is cut into contiguous,
window-aligned row ranges of (nearly) equal weight -- nnz by default, TC blocks if the caller has them --
one per rank.  Each rank preprocesses only its shard.  The only exchange step of the path is making B
available everywhere: one NCCL broadcast (B is needed in full by every shard, because any shard may
reference any column).  C stays row-sharded; ``all_gather`` materialises it on every rank on request
(shards are padded to the largest one, NCCL all_gather needs equal sizes).

Window alignment makes shard windows coincide with the single-GPU windows, so the concatenated shard
outputs are bit-identical to the 1-GPU result.

One process per GPU, ``torch.distributed`` for the plumbing (backend nccl on GPUs; the host logic is
covered with gloo on CPU in tests/test_distributed.py, where the local SpMM is injected).
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

BLK_H = 16
ROW_COST = 7    # in units of one non-zero, see ShardedSpMM.__init__


def partition_rows(weights: torch.Tensor, world_size: int, align: int = BLK_H) -> List[Tuple[int, int]]:
    """Contiguous, ``align``-aligned row ranges minimising the heaviest shard (bottleneck-optimal).

    ``weights[r]`` is the cost of row r (its nnz, or its share of the window's TC blocks).  The smallest
    per-shard budget L for which a greedy left-to-right packing of whole windows fits in ``world_size``
    shards is found by bisection (each probe is ``world_size`` binary searches on the window prefix sums),
    then that packing is returned.  A hub window heavier than the ideal share becomes a shard of its own;
    trailing shards may be empty when there are fewer windows than ranks.  Deterministic: every rank
    computes the same answer from the same input.
    """
    M = int(weights.numel())
    W = (M + align - 1) // align
    if W == 0:
        return [(0, 0)] * world_size
    w = weights.to(torch.int64).cpu()
    pad = W * align - M
    if pad:
        w = torch.cat([w, torch.zeros(pad, dtype=w.dtype)])
    ww = w.view(W, align).sum(1)
    prefix = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(ww, 0)])   # prefix[i] = weight of windows [0, i)
    total, heaviest = int(prefix[-1]), int(ww.max())

    def pack(limit: int):
        cuts, start = [0], 0
        for _ in range(world_size):
            if start >= W:
                cuts.append(W)
                continue
            # furthest end with prefix[end] - prefix[start] <= limit
            end = int(torch.searchsorted(prefix, prefix[start] + limit, right=True)) - 1
            end = max(end, start + 1) if limit >= int(ww[start]) else start   # a window over budget: infeasible
            if end == start:
                return None
            cuts.append(min(end, W))
            start = cuts[-1]
        return cuts if start >= W else None

    lo, hi = max(heaviest, (total + world_size - 1) // world_size), max(total, 1)
    while lo < hi:
        mid = (lo + hi) // 2
        if pack(mid) is not None:
            hi = mid
        else:
            lo = mid + 1
    cuts = pack(lo)
    return [(min(cuts[k] * align, M), min(cuts[k + 1] * align, M)) for k in range(world_size)]


def shard_csr(indptr: torch.Tensor, indices: torch.Tensor, r0: int, r1: int) -> Tuple[torch.Tensor, torch.Tensor]:
    lo, hi = int(indptr[r0]), int(indptr[r1])
    return (indptr[r0:r1 + 1] - lo).to(torch.int32).contiguous(), indices[lo:hi].contiguous()


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (``/sys/bus/pci/devices/<bus id>/numa_node``),
    so that pinned host buffers allocated afterwards (first touch) live next to the GPU's PCIe root: host<->device copies of
    the end-to-end path then do not cross the socket interconnect.  Returns the node, or None when the topology is not
    exposed (single-socket hosts, containers without sysfs)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, AttributeError, ValueError, RuntimeError, AssertionError):
        return None


def operand_slices(rows: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Equal row slices of the dense operand for the per-step exchange: ``chunk = ceil(rows / world)`` rows per rank (what
    an all-gather needs), the last ranks' slices clipped to ``rows`` (possibly empty).  Returns ``(chunk, [(lo, hi)])``."""
    chunk = (rows + world - 1) // world
    return chunk, [(min(r * chunk, rows), min((r + 1) * chunk, rows)) for r in range(world)]


def upload_slice_and_all_gather(buf: torch.Tensor, feat_host: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """The per-step exchange of a host-resident dense operand (north_star: "B is broadcast over NVLink with NCCL"):
    every rank copies only ITS 1/world row slice of ``feat_host`` (pinned host memory on a GPU box) into ``buf`` and the
    ranks all-gather the slices in place, so each step moves the operand ONCE over the host links (1/world per GPU) and
    the rest over NVLink.  ``buf``: ``[chunk * world, N]`` on the compute device; returns ``buf[:rows]``.  Runs on the
    current stream (the NCCL work is ordered after the copy and before whatever the caller enqueues next)."""
    rows = feat_host.shape[0]
    chunk, slices = operand_slices(rows, world)
    assert buf.shape[0] == chunk * world and buf.shape[1:] == feat_host.shape[1:]
    lo, hi = slices[rank]
    if hi > lo:
        buf[lo:hi].copy_(feat_host[lo:hi], non_blocking=True)
    if world > 1:
        mine = buf[rank * chunk:(rank + 1) * chunk]
        if buf.is_cuda:
            dist.all_gather_into_tensor(buf, mine, group=group)      # in place: `mine` is buf's own slot of this rank
        else:
            parts = [torch.empty_like(mine) for _ in range(world)]   # gloo (CPU tests): no in-place flat gather
            dist.all_gather(parts, mine.clone(), group=group)
            for r, part in enumerate(parts):
                buf[r * chunk:(r + 1) * chunk].copy_(part)
    return buf[:rows]


class ShardedSpMM:
    """Per-rank state of a row-sharded SpMM.

    ``local_preprocess(indptr, indices, num_rows, num_cols)`` and ``local_spmm(state, feat) -> C_local`` default
    to the CUDA path (``voltrix.csr_preprocess`` / ``voltrix.spmm``) and raise without a GPU; tests inject CPU
    stand-ins to exercise the partitioning / collective logic under gloo.
    """

    def __init__(self, indptr: Optional[torch.Tensor], indices: Optional[torch.Tensor], num_nodes: int,
                 weights: Optional[torch.Tensor] = None, group=None,
                 local_preprocess: Optional[Callable] = None, local_spmm: Optional[Callable] = None,
                 local_csr: Optional[Callable] = None):
        """Either the whole CSR matrix (``indptr``, ``indices``; every rank holds it and cuts out its shard), or -- for
        graphs no single rank should materialise -- per-row ``weights`` for the partition plus ``local_csr(r0, r1) ->
        (indptr, indices)``, which builds just this rank's rows (column ids global)."""
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.num_nodes = num_nodes
        if weights is None:
            # cost of a row ~ its non-zeros (one gathered B row each) + a fixed share of its window's work item
            # (schedule slot, TMEM epilogue, the C row write).  Fitted on the per-rank times of the 8-GPU R-MAT-25 run
            # (N = 256; profiles/r2z_bench_n8.json and its history in DESIGN.md section 5): t = 0.039 ms per 10^6 non-zeros
            # + 0.257 ms per 10^6 rows => one row costs 6.6 non-zeros.  Round 1 used 12 (fitted at scale 21 on the one-CTA
            # kernel, before the CUDA-core rows took four rows per warp), which left the row-heavy last rank 15 % short of
            # work at 8 GPUs.
            weights = (indptr[1:] - indptr[:-1]) + ROW_COST
        self.ranges = partition_rows(weights, self.world)
        self.r0, self.r1 = self.ranges[self.rank]
        self.local_rows = self.r1 - self.r0
        if local_csr is not None:
            lp, li = local_csr(self.r0, self.r1)
            assert lp.numel() == self.local_rows + 1
        else:
            lp, li = shard_csr(indptr, indices, self.r0, self.r1)
        self.local_nnz = int(li.numel())
        if local_preprocess is None:
            from .spmm import csr_preprocess
            local_preprocess = lambda ip, ix, rows, cols: csr_preprocess(ip, ix, rows, num_cols=cols)  # noqa: E731
        if local_spmm is None:
            from .spmm import spmm as _spmm
            local_spmm = lambda st, feat: _spmm(st[0], st[1], st[2], self.local_rows, self.local_nnz, feat)  # noqa: E731
        self._spmm = local_spmm
        self.state = local_preprocess(lp, li, self.local_rows, num_nodes) if self.local_rows > 0 else None

    # -- the one exchange step of the path --------------------------------------------------------
    def broadcast_features(self, feat: torch.Tensor, src: int = 0) -> torch.Tensor:
        """Make the full dense operand resident on every rank (NCCL broadcast over NVLink/NVSwitch)."""
        if self.world > 1:
            dist.broadcast(feat, src=src, group=self.group)
        return feat

    def spmm(self, feat: torch.Tensor) -> torch.Tensor:
        """C[r0:r1, :] for this rank; ``feat`` is the FULL dense operand, already resident."""
        if self.local_rows == 0:
            return torch.empty((0, feat.shape[1]), dtype=torch.float32, device=feat.device)
        return self._spmm(self.state, feat)

    def all_gather(self, c_local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Optional: full C ``[num_nodes, N]`` on every rank.  Every rank sends exactly its own rows (grouped
        point-to-point sends inside one NCCL group call -- no padding to the largest shard and no concatenation pass);
        ``out`` lets the caller reuse the result buffer across steps."""
        if self.world == 1:
            return c_local
        n = c_local.shape[1]
        if out is None:
            out = torch.empty((self.num_nodes, n), dtype=c_local.dtype, device=c_local.device)
        parts = [out[a:b] for a, b in self.ranges]
        peer = (lambda k: dist.get_global_rank(self.group, k)) if self.group is not None else (lambda k: k)
        if c_local.is_cuda:
            # uneven all-gather: NCCL runs the world x world sends / receives as one fused group
            ops = []
            for k in range(self.world):
                if k == self.rank:
                    continue
                if self.local_rows > 0:
                    ops.append(dist.P2POp(dist.isend, c_local, peer(k), self.group))
                if parts[k].shape[0] > 0:
                    ops.append(dist.P2POp(dist.irecv, parts[k], peer(k), self.group))
            parts[self.rank].copy_(c_local)
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        else:
            for k in range(self.world):   # gloo (CPU tests): one broadcast per shard
                if parts[k].shape[0] == 0:
                    continue
                if k == self.rank:
                    parts[k].copy_(c_local)
                dist.broadcast(parts[k], src=peer(k), group=self.group)
        return out

    def imbalance(self, weights: Sequence[float]) -> float:
        """max / mean of per-rank weights (1.0 = perfect)."""
        w = list(weights)
        return max(w) / (sum(w) / len(w)) if sum(w) else 1.0
