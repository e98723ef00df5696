"""Where the time of a 1/8 shard of the Reddit-shaped SpMM goes (the per-rank launch of the 8-GPU run, replayed on one GPU):
CUDA-event time per step with and without the L2 flush (cold vs warm dense operand), and the kernel's own duration from the
profiler -- the difference is launch gaps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
import bench as B  # noqa: E402
from voltrix.distributed import ROW_COST, partition_rows, shard_csr  # noqa: E402

dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload("reddit", dev, 1.0)
M = indptr.numel() - 1
feat = torch.rand(M, N, device=dev).half()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
full = None
for world in (1, 2, 4, 8):
    r0, r1 = partition_rows((indptr[1:] - indptr[:-1]) + ROW_COST, world)[0]
    lp, li = shard_csr(indptr, indices, r0, r1)
    rows, nnz = r1 - r0, li.numel()
    st = voltrix.csr_preprocess(lp, li, rows, num_cols=M)
    out = torch.empty(rows, N, device=dev)
    fn = lambda: voltrix.spmm(*st, rows, nnz, feat, out=out)   # noqa: E731
    fn(); torch.cuda.synchronize()

    def timed(do_flush, iters=20):
        ts = []
        for _ in range(iters):
            if do_flush:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts))

    cold, warm = timed(True), timed(False)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            flush.zero_(); fn()
        torch.cuda.synchronize()
    k = {e.key[:40]: e.device_time_total / max(e.count, 1) / 1e3 for e in prof.key_averages() if "spmm" in e.key}
    if full is None:
        full = cold
    print(f"shard 1/{world}: rows {rows} nnz {nnz} items {st[1]._vx_plan.num_items} | events cold-L2 {cold:.4f} ms "
          f"(ideal {full / world:.4f}, eff {full / world / cold:.3f}) | warm-L2 {warm:.4f} ms | kernels (profiler, cold) "
          f"{ {n: round(v, 4) for n, v in k.items()} }", flush=True)
