"""Why do bench/bm_voltrix.py's kineto times differ from CUDA-event times?  Same call, three clocks."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
from voltrix import graphs  # noqa: E402
from voltrix.utils import GPU_bench  # noqa: E402

dev = torch.device("cuda")
for name, N, dt in (("amazon0505", 256, torch.float16), ("amazon0505", 256, torch.float32), ("FraudYelp-RSR", 256, torch.float16)):
    indptr, indices = graphs.suite_graph(name, seed=0, device=dev)
    M, nnz = indptr.numel() - 1, indices.numel()
    st = voltrix.csr_preprocess(indptr.cpu(), indices.cpu(), M)
    feat = torch.rand(M, N, device=dev).to(dt)
    fn = lambda: voltrix.spmm(*st, M, nnz, feat)   # noqa: E731
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    kin = GPU_bench(fn, iters=10, warmup=10, kernel_name="spmm")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            flush.zero_(); fn()
        torch.cuda.synchronize()
    rows = [(e.key[:70], e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if "spmm" in e.key or "Memset" in e.key]
    print(f"{name} N={N} {dt}: events median {np.median(ts):.4f} ms | GPU_bench(kernel_name='spmm') {kin:.4f} ms | tuned "
          f"{[v for k, v in voltrix.jit_tuner.tuned_keys.items() if f'{N},' in k[1]][-1:]}")
    for r in rows:
        print("     ", r)
