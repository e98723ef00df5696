"""Per-edge values: the WEIGHTED tensor-core kernel beside the binary kernel and the weighted CUDA-core rows on one workload.
    python scripts/weighted_probe.py [--workload reddit] [--scale 1.0]"""
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
args = ap.parse_args()
dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload(args.workload, dev, args.scale)
M, nnz = indptr.numel() - 1, indices.numel()
st = voltrix.csr_preprocess(indptr, indices, M)
vals = torch.rand(nnz, device=dev) + 0.5
w = voltrix.edge_weights(*st, indptr, indices, vals)
feat = torch.rand(M, N, device=dev).half()
out = torch.empty(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
print(f"{desc}: M={M} nnz={nnz} N={N} TCB={st[1]._vx_plan.total_blocks}")


def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def variant(model, stages, ew):
    return lambda: voltrix.spmm_kernel(*st, num_nodes=M, num_edges=nnz, embedding_dim=N, input=feat, output=out, model=model,
                                       stages=stages, edge_weights=ew)


t_tiles = timeit(lambda: w._tiles.clear() or w.tiles(torch.float16))
print(f"value tiles build: {t_tiles:.3f} ms")
for name, fn in (("binary 14/7", variant(0, 14, None)), ("weighted tc 14/7", variant(0, 14, w)),
                 ("weighted tc 22/11", variant(0, 22, w)), ("weighted tc 42/14", variant(0, 42, w)),
                 ("weighted cuda-core rows (model 1)", variant(1, 32, w)),
                 ("weighted autotuned", lambda: voltrix.spmm(*st, M, nnz, feat, out=out, edge_weights=w))):
    ms = timeit(fn)
    print(f"{name:36s} {ms:8.3f} ms  {2.0 * nnz * N / ms / 1e6:9.1f} GFLOP/s")
# value check against torch.sparse on a row sample
rows = min(M, 20000)
lo, hi = 0, int(indptr[rows])
A = torch.sparse_csr_tensor(indptr[: rows + 1], indices[:hi], vals[:hi], size=(rows, M))
want = A @ feat.float()
voltrix.spmm(*st, M, nnz, feat, out=out, edge_weights=w)
err = ((out[:rows] - want).abs().max() / want.abs().max()).item()
print(f"weighted autotuned vs torch.sparse fp32 on rows [0,{rows}): max scaled err {err:.2e}")
