"""voltrix.tune_routing on C3 shapes: the routing rule csr_preprocess installs (sparse_ratio 0.5; small_blocks 8 from 4 M TC blocks up, else 0) against the measured best of
voltrix.ROUTING_CANDIDATES, per graph / width / dtype.

    python scripts/routing_probe.py [--datasets ppi protein ...] [--feature_dims 64 128 256] [--out gpurun_out/routing.csv]"""
import argparse
import csv
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
from voltrix import graphs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--datasets", nargs="*", default=["ppi", "protein", "DD", "com-amazon", "amazon0505", "web-BerkStan", "FraudYelp-RSR"])
ap.add_argument("--feature_dims", nargs="*", type=int, default=[64, 128, 256])
ap.add_argument("--out", default="gpurun_out/routing.csv")
args = ap.parse_args()
dev = torch.device("cuda")
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["dataset", "N", "dtype", "default_ms", "best_rule", "best_ms", "gain", "all_rules_ms"])
    for name in args.datasets:
        indptr, indices = graphs.suite_graph(name, seed=0, device=dev)
        M, nnz = indptr.numel() - 1, indices.numel()
        st = voltrix.csr_preprocess(indptr, indices, M)
        for N in args.feature_dims:
            for dtype in (torch.float16, torch.float32):
                feat = torch.rand(M, N, device=dev).to(dtype)
                want = voltrix.spmm(*st, M, nnz, feat)
                installed = (st[1]._vx_plan.sparse_ratio, st[1]._vx_plan.small_blocks)
                best, timings = voltrix.tune_routing(*st, M, nnz, feat)
                got = voltrix.spmm(*st, M, nnz, feat)
                err = float((got - want).abs().max() / want.abs().max().clamp_min(1e-9))
                default = timings[installed]
                row = [name, N, str(dtype)[6:], f"{default:.4f}", f"{best[0]:g}/{best[1]}", f"{timings[best]:.4f}",
                       f"{default / timings[best]:.3f}", " ".join(f"{k[0]:g}/{k[1]}:{v:.4f}" for k, v in timings.items())]
                assert err < 2e-3, (name, N, dtype, err)
                w.writerow(row); fh.flush()
                print(" ".join(str(x) for x in row), flush=True)
                voltrix.reschedule(*st, *installed)
