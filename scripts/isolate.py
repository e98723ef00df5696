"""Bottleneck isolation of the tensor-core kernel: times the normal build and the VX_TC_DBG=1 (no MMA) /
VX_TC_DBG=2 (no TMA) builds of one variant on one workload.  The debug builds compute garbage; timing only.
Prebuild them first on the build box: VOLTRIX_EXTRA_NVCC_FLAGS=-DVX_TC_DBG=1 python scripts/prebuild_variants.py 0/36/12"""
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
from voltrix.jit_kernels.tuner import jit_tuner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="reddit")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--variant", default="0/36/12")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--modes", default="0,1,2")
ap.add_argument("--flags", default=None, help="'|'-separated VOLTRIX_EXTRA_NVCC_FLAGS values to compare, e.g. --flags=\"|-DX=1|-DX=2\" (overrides --modes)")
args = ap.parse_args()
dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload(args.workload, dev, args.scale)
M, nnz = indptr.numel() - 1, indices.numel()
blk, packed, hind = voltrix.csr_preprocess(indptr, indices, M)
plan = packed._vx_plan
feat = torch.rand(M, N, device=dev).half()
out = torch.empty(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
model, stages, npw = (int(x) for x in args.variant.split("/"))
print(f"{desc}: M={M} nnz={nnz} N={N} TCB={plan.total_blocks}")
flag_sets = args.flags.split("|") if args.flags is not None else ["" if m == "0" else f"-DVX_TC_DBG={m}" for m in args.modes.split(",")]
ref = None
for mode in flag_sets:
    if mode == "":
        os.environ.pop("VOLTRIX_EXTRA_NVCC_FLAGS", None)
    else:
        os.environ["VOLTRIX_EXTRA_NVCC_FLAGS"] = mode
    jit_tuner.tuned.clear()
    getattr(packed, "_vx_fast", {}).clear()      # the prepared launch of the previous flag set

    def run():
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=nnz, embedding_dim=N, input=feat, output=out,
                            model=model, stages=stages, npw=npw)
    run(); torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"flags '{mode}' variant {args.variant}: median {np.median(ts):.3f} ms (min {min(ts):.3f} max {max(ts):.3f}) "
          f"max scaled diff vs first {err:.2e}")
