"""cuSPARSE baseline beside bench/bm_voltrix.py: `torch.sparse_csr @ dense` on the files graph_gen.py wrote into the CWD.
Prints the three lines bench_all.py and the reference's README read: nnz, the allclose verdict against output_base.csv
(atol 1e-1, the reference's bar, bench/bm_sparse.py:50) and `[cuSPARSE] Elapsed time: x ms` (10 warm-up + 100 timed
back-to-back calls between CUDA events -- voltrix.utils.GPU_bench without a kernel name)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "voltrix-spmm_b200")))
from voltrix.utils import GPU_bench  # noqa: E402


def load_csv_int32(name):
    return torch.from_numpy(np.loadtxt(name, delimiter=",", dtype=np.int32)).cuda()


col, rowptr = load_csv_int32("indices.csv"), load_csv_int32("indptr.csv")
rows = rowptr.numel() - 1
dense = torch.from_numpy(np.fromfile("feat.csv", dtype=np.float32)).cuda().view(rows, -1)
adjacency = torch.sparse_csr_tensor(rowptr, col, torch.ones(col.numel(), device="cuda"), size=(rows, rows))
print(col.numel())
ms = GPU_bench(lambda: adjacency @ dense, iters=100, warmup=10)
expected = np.fromfile("output_base.csv", dtype=np.float32).reshape(rows, -1)
print(np.allclose((adjacency @ dense).cpu().numpy(), expected, atol=1e-1))
print(f"[cuSPARSE] Elapsed time: {ms:.4f} ms")
