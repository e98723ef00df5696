"""CUDA-core CSR path on the sparsest C3 shape (YeastH: mean degree ~2): a few launches of voltrix.spmm at one width / dtype
for an ncu capture (`--set full -k regex:csr`), and the CUDA-event time beside the compulsory and the no-reuse DRAM bytes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
import bench as B  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "YeastH"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dtype = {"fp16": torch.float16, "fp32": torch.float32}[sys.argv[3] if len(sys.argv) > 3 else "fp16"]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda")
indptr, indices, _, desc = B.make_workload(name, dev, 1.0)
M, nnz = indptr.numel() - 1, indices.numel()
feat = torch.rand(M, N, device=dev).to(dtype)
st = voltrix.csr_preprocess(indptr, indices, M)
out = torch.empty(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
fn = lambda: voltrix.spmm(*st, M, nnz, feat, out=out)   # noqa: E731
fn(); torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
t = float(np.median(ts))
es = feat.element_size()
alg = 4 * nnz + 4 * (M + 1) + M * N * es + M * N * 4
noreuse = 4 * nnz + 4 * (M + 1) + nnz * N * es + M * N * 4
plan = st[1]._vx_plan
print(f"{name} M={M} nnz={nnz} N={N} {dtype}: {t:.4f} ms | ALG {alg / 1e9:.3f} GB -> {alg / t / 1e6:.0f} GB/s | every gathered row "
      f"from DRAM {noreuse / 1e9:.3f} GB -> {noreuse / t / 1e6:.0f} GB/s | sparse rows {plan.num_sparse_rows}/{M} items {plan.num_items}")
