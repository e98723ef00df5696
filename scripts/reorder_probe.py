"""Measured effect of voltrix.reorder.lsh_reorder / cluster_reorder on a graph with planted communities and shuffled labels:
TC blocks and SpMM time before / after relabelling, result checked through the inverse permutation."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
from voltrix import reorder  # noqa: E402
from voltrix.graphs import planted_partition_csr  # noqa: E402

dev = torch.device("cuda")
M, N = 262144, 128
indptr, indices = planted_partition_csr(M, community=256, p_in=0.2, p_out=2e-6, seed=0, device=dev)
nnz = indices.numel()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
feat = torch.rand(M, N, device=dev).half()


def timeit(fn, iters=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


st = voltrix.csr_preprocess(indptr, indices, M)
t0 = timeit(lambda: voltrix.spmm(*st, M, nnz, feat))
base = voltrix.spmm(*st, M, nnz, feat)
print(f"planted partition M={M} nnz={nnz} N={N} fp16")
print(f"shuffled labels : TCB={st[1]._vx_plan.total_blocks:9d}  spmm {t0:.3f} ms")
for name, fn in (("lsh_reorder", reorder.lsh_reorder), ("cluster_reorder", reorder.cluster_reorder)):
    torch.cuda.synchronize(); t = time.perf_counter()
    perm = fn(indptr, indices)
    ip2, ix2 = reorder.permute_graph(indptr, indices, perm)
    torch.cuda.synchronize(); t_re = time.perf_counter() - t
    st2 = voltrix.csr_preprocess(ip2, ix2, M)
    feat2 = reorder.permute_rows(feat, perm)
    t1 = timeit(lambda: voltrix.spmm(*st2, M, nnz, feat2))
    got = reorder.unpermute_rows(voltrix.spmm(*st2, M, nnz, feat2), perm)
    err = ((got - base).abs().max() / base.abs().max()).item()
    print(f"{name:16s}: TCB={st2[1]._vx_plan.total_blocks:9d}  spmm {t1:.3f} ms  ({t0 / t1:.2f}x)  reorder+relabel {t_re * 1e3:.0f} ms  "
          f"max scaled diff after un-permuting {err:.1e}")
ip0, ix0 = planted_partition_csr(M, community=256, p_in=0.2, p_out=2e-6, seed=0, device=dev, shuffle=False)
st0 = voltrix.csr_preprocess(ip0, ix0, M)
t_ideal = timeit(lambda: voltrix.spmm(*st0, M, ix0.numel(), feat))
print(f"unshuffled graph: TCB={st0[1]._vx_plan.total_blocks:9d}  spmm {t_ideal:.3f} ms  (the order the generator planted)")
