"""DTC-SpMM baseline (reference: bench/bm_dtc.py; kernels: third-party/DTC-SpMM/DTC-SpMM).

Loads the `DTCSpMM` torch extension bench/competitors/build.py compiled for sm_100a, runs its GPU preprocessing
(`preprocess_gpu`) and its `run_DTCSpMM` entry point with the execution plan the reference script uses (`float4_split`, or the
balanced variant with --use_balance) on the graph_gen.py files in the CWD.  The extension times 1000 back-to-back launches
itself and appends `<name>,<ms>,<GFLOP/s>` to DTCSpMM_exe_time_and_throughput.csv; that figure is what bench_all.py parses."""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_competitors", "dtc"))
try:
    import DTCSpMM  # noqa: E402
except ImportError as e:  # pragma: no cover
    raise SystemExit(f"DTCSpMM extension not found ({e}) -- run `python bench/competitors/build.py` first")

BLK_H, BLK_W = 16, 8
ap = argparse.ArgumentParser()
ap.add_argument("--use_balance", action="store_true", help="the row-window-balanced kernel (reference flag)")
args = ap.parse_args()

col = torch.from_numpy(np.loadtxt("indices.csv", delimiter=",", dtype=np.int32)).cuda()
rowptr = torch.from_numpy(np.loadtxt("indptr.csv", delimiter=",", dtype=np.int32)).cuda()
rows, nnz = rowptr.numel() - 1, col.numel()
dense = torch.from_numpy(np.fromfile("feat.csv", dtype=np.float32)).cuda().view(rows, -1)
windows = (rows + BLK_H - 1) // BLK_H
block_partition = torch.zeros(windows, dtype=torch.int32, device="cuda")
edge_to_col = torch.zeros(nnz, dtype=torch.int32, device="cuda")
edge_to_row = torch.zeros(nnz, dtype=torch.int32, device="cuda")
win_off, blk_row, tile_id, blk_off, a_to_x, _ = DTCSpMM.preprocess_gpu(col, rowptr, rows, BLK_H, BLK_W, block_partition,
                                                                     edge_to_col, edge_to_row)
log = "DTCSpMM_exe_time_and_throughput.csv"
if os.path.exists(log):
    os.remove(log)
torch.cuda.synchronize()
if args.use_balance:
    out = DTCSpMM.run_DTCSpMM_balance(dense, blk_row, tile_id, blk_off, a_to_x, rows, "float4_split")[0]
else:
    out = DTCSpMM.run_DTCSpMM(dense, win_off, tile_id, blk_off, a_to_x, rows, nnz, "float4_split")[0]
torch.cuda.synchronize()
expected = np.fromfile("output_base.csv", dtype=np.float32).reshape(rows, -1)
print(np.allclose(out.cpu().numpy(), expected, atol=1e-1))
with open(log) as fh:
    print(f"[DTC-SPMM] Elapsed time: {float(fh.readline().split(',')[1]):.4f} ms")
