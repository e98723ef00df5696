"""Times every SpMM kernel variant on a workload (L2 flushed between iterations) and prints a table.
    python scripts/time_models.py [--workload reddit] [--scale 1.0] [--iters 10] [--only MODEL/STAGES] [--once]
--once: a single launch of the selected variant (for ncu captures)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
import bench as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="reddit")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--only", default=None)
ap.add_argument("--once", action="store_true")
ap.add_argument("--sparse_ratio", type=float, default=None)
ap.add_argument("--small_blocks", type=int, default=None)
ap.add_argument("--N", type=int, default=None)
args = ap.parse_args()

dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload(args.workload, dev, args.scale)
if args.N:
    N = args.N
M, nnz = indptr.numel() - 1, indices.numel()
dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
kw = {} if args.sparse_ratio is None else {"sparse_ratio": args.sparse_ratio}
if args.small_blocks is not None:
    kw["small_blocks"] = args.small_blocks
blk, packed, hind = voltrix.csr_preprocess(indptr, indices, M, **kw)
plan = packed._vx_plan
print(f"{desc}: M={M} nnz={nnz} N={N} TCB={plan.total_blocks} items={plan.num_items} sparse_rows={plan.num_sparse_rows} "
      f"fixups={plan.num_fixups} cap={plan.cap}")
feat = torch.rand(M, N, device=dev).to(dt)
out = torch.empty(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
variants = [(0, 16, 4), (0, 32, 8), (0, 36, 12), (1, 32, 8), (2, 32, 8)] if dt != torch.float32 else [(3, 24, 8), (1, 32, 8), (2, 32, 8)]
if args.only:
    variants = [tuple(int(x) for x in v.split("/")) for v in args.only.split(",")]
ref = None
for model, stages, npw, *rest in variants:
    ft = rest[0] if rest else None

    def run():
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=nnz, embedding_dim=N, input=feat, output=out,
                            model=model, stages=stages, npw=npw, ft=ft)
    if args.once:
        run(); torch.cuda.synchronize(); flush.zero_(); run(); torch.cuda.synchronize()
        print(f"ran model {model} stages {stages} npw {npw} once (after one warm-up)")
        continue
    run(); torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    else:
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        assert err < 5e-3, (model, stages, err)   # hub rows of an R-MAT sum millions of terms: summation order shows
        print(f"   (max scaled diff vs first variant {err:.2e})")
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    gather = plan.total_blocks * 8 * N * feat.element_size() if model != 1 else nnz * N * feat.element_size()
    print(f"model {model} stages {stages:2d} npw {npw:2d} ft {ft or 128:3d}: {ms:8.3f} ms  {2.0 * nnz * N / ms / 1e6:9.1f} GFLOP/s  "
          f"gather {gather / ms / 1e6:8.1f} GB/s  (min {min(ts):.3f} max {max(ts):.3f})")
