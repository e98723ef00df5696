"""C3 (BASELINE.json configs[2]): the GNN / SuiteSparse-shaped suite x N sweep in ONE process -- voltrix.spmm (autotuned,
fp16 and fp32) beside cuSPARSE (torch.sparse_csr @ dense, the reference's bench/bm_sparse.py protocol) on the same inputs.
bench/bench_all.py is the reference-shaped driver (one subprocess per method); this script produces the same table without
paying ~10 s of interpreter start-up per cell.

    python scripts/suite.py [--datasets ddi ppi ...] [--feature_dims 32 64 128 256 512] [--out gpurun_out/suite.csv]
Timing: L2 flushed (256 MB write) before every launch, CUDA events, median of --iters."""
import argparse
import csv
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
from voltrix import graphs  # noqa: E402
from voltrix.utils import calc_diff  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--datasets", nargs="*", default=None)
ap.add_argument("--feature_dims", nargs="*", type=int, default=[32, 64, 128, 256, 512])
ap.add_argument("--iters", type=int, default=7)
ap.add_argument("--out", default="gpurun_out/suite.csv")
args = ap.parse_args()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)


def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


names = args.datasets or [n for n, _, _ in graphs.named_suite()]
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["dataset", "M", "nnz", "TCB", "N", "cusparse_fp32_ms", "cusparse_fp16_ms", "voltrix_fp32_ms", "voltrix_fp16_ms",
                "fp16_speedup_vs_cusparse_fp16", "fp32_speedup_vs_cusparse_fp32", "voltrix_fp16_gflops", "fp16_diff_rate_pct",
                "tuned_fp16"])
    for name in names:
        indptr, indices = graphs.suite_graph(name, seed=0, device=dev)
        M, nnz = indptr.numel() - 1, indices.numel()
        blk, packed, hind = voltrix.csr_preprocess(indptr, indices, M)
        packed.hash_tag = f"suite-{name}"
        tcb = packed._vx_plan.total_blocks
        csr32 = torch.sparse_csr_tensor(indptr, indices, torch.ones(nnz, device=dev), size=(M, M))
        csr16 = torch.sparse_csr_tensor(indptr, indices, torch.ones(nnz, device=dev, dtype=torch.float16), size=(M, M))
        for N in args.feature_dims:
            f32 = torch.rand(M, N, device=dev)
            f16 = f32.half()
            t_c32 = timeit(lambda: csr32 @ f32)
            try:
                t_c16 = timeit(lambda: csr16 @ f16)
            except Exception:
                t_c16 = float("nan")
            t_v32 = timeit(lambda: voltrix.spmm(blk, packed, hind, M, nnz, f32))
            t_v16 = timeit(lambda: voltrix.spmm(blk, packed, hind, M, nnz, f16))
            diff = calc_diff(voltrix.spmm(blk, packed, hind, M, nnz, f16).cpu(), (csr32 @ f16.float()).cpu()) * 100
            tuned = [v for k, v in voltrix.jit_tuner.tuned_keys.items()
                     if k[0] == "spmm_kernel" and f"'N': {N}," in k[1] and "__half" in k[1]]
            row = [name, M, nnz, tcb, N, f"{t_c32:.4f}", f"{t_c16:.4f}", f"{t_v32:.4f}", f"{t_v16:.4f}",
                   f"{t_c16 / t_v16:.2f}", f"{t_c32 / t_v32:.2f}", f"{2.0 * nnz * N / t_v16 / 1e6:.0f}", f"{diff:.3f}",
                   "/".join(str(tuned[-1][k]) for k in ("model", "stages", "npw")) if tuned else "?"]
            w.writerow(row); fh.flush()
            print(" ".join(str(x) for x in row), flush=True)
        del csr32, csr16, blk, packed, hind
        torch.cuda.empty_cache()
