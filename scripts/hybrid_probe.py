"""Probe: tensor-core kernel on rows [0, r) and the CUDA-core CSR kernel on rows [r, M) CONCURRENTLY (two streams),
for several split fractions.  Question: does the CUDA-core kernel, which needs no shared-memory staging and no MMA,
soak up the L2 bandwidth the tensor-core kernel leaves unused (lts 65 %)?"""
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
from voltrix.distributed import shard_csr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="reddit")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--fracs", default="0,0.1,0.15,0.2,0.25,0.3,0.35")
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--variant", default="0/36/12")
args = ap.parse_args()
dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload(args.workload, dev, args.scale)
M, nnz = indptr.numel() - 1, indices.numel()
feat = torch.rand(M, N, device=dev).half()
out = torch.empty(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
model, stages, npw = (int(x) for x in args.variant.split("/"))
s_tc, s_cc = torch.cuda.Stream(), torch.cuda.Stream()
print(f"{desc}: M={M} nnz={nnz} N={N}")
ref = None
for frac in (float(x) for x in args.fracs.split(",")):
    # rows [0, r) -> tensor cores, [r, M) -> CUDA cores; r window-aligned, split by nnz
    target = int(nnz * (1.0 - frac))
    r = int(torch.searchsorted(indptr.long(), torch.tensor([target], device=dev)).item())
    r = min(M, (r + 15) // 16 * 16)
    parts = []
    for (a, b, mdl) in ((0, r, model), (r, M, 1)):
        if b <= a:
            parts.append(None)
            continue
        lp, li = shard_csr(indptr, indices, a, b)
        st = voltrix.csr_preprocess(lp, li, b - a, num_cols=M, sparse_ratio=0.0)
        parts.append((st, b - a, li.numel(), out[a:b], mdl))

    def launch(p):
        (blk, packed, hind), rows, e, o, mdl = p
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=rows, num_edges=e, embedding_dim=N, input=feat, output=o,
                            model=mdl, stages=stages if mdl == 0 else 32, npw=npw if mdl == 0 else 8)

    def run():
        cur = torch.cuda.current_stream()
        s_tc.wait_stream(cur); s_cc.wait_stream(cur)
        if parts[0] is not None:
            with torch.cuda.stream(s_tc):
                launch(parts[0])
        if parts[1] is not None:
            with torch.cuda.stream(s_cc):
                launch(parts[1])
        cur.wait_stream(s_tc); cur.wait_stream(s_cc)

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
    print(f"cuda-core fraction {frac:.2f} (rows [{r},{M})): median {np.median(ts):.3f} ms (min {min(ts):.3f}) diff {err:.1e}")
