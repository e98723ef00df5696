"""A small end-to-end run of every SpMM path for compute-sanitizer (memcheck / synccheck / initcheck): preprocessing, the
tcgen05 kernel in its shipped shapes (128- and 64-wide feature tile, K-split fix-up, weighted, fp32 as one / two terms), and
the CUDA-core kernels, each checked against scipy so a silent mis-read would also show up as a wrong result."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402

rng = np.random.default_rng(7)
M = 3000 + 5                                      # a partial tail window
A = sp.random(M, M, density=0.004, format="csr", random_state=3, dtype=np.float32)
hub = sp.random(16, M, density=0.9, format="csr", random_state=4, dtype=np.float32)      # one dense window: K-split + fix-up
A = sp.vstack([hub, A[16:]]).tocsr(); A.sum_duplicates(); A.sort_indices()
indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
vals = rng.uniform(0.5, 1.5, indices.size).astype(np.float32)
ones = sp.csr_matrix((np.ones(indices.size, np.float32), indices, indptr), shape=(M, M))
wtd = sp.csr_matrix((vals, indices, indptr), shape=(M, M))
st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
plan = st[1]._vx_plan
print(f"M={M} nnz={indices.size} items={plan.num_items} fixups={plan.num_fixups} sparse_rows={plan.num_sparse_rows} cap={plan.cap}")
bad = 0


def check(tag, got, want, tol):
    global bad
    err = float(np.abs(got - want).max() / max(1.0, np.abs(want).max()))
    ok = np.isfinite(got).all() and err <= tol
    bad += not ok
    print(f"{'ok ' if ok else 'BAD'} {tag}: max scaled err {err:.2e}")


for N in (128, 64, 32, 200):
    B = rng.standard_normal((M, N)).astype(np.float32)
    want = ones @ B
    for dtype, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1.5e-2), (torch.float32, 1e-3)):
        x = torch.from_numpy(B).cuda().to(dtype)
        ref = ones @ x.float().cpu().numpy()
        models = [None, 0, 1, 2] if dtype != torch.float32 else [None, 1, 2, 3, 4]
        for model in models:
            if model in (0, 3, 4) and N % 8:
                continue
            out = torch.full((M, N), float("nan"), device="cuda")
            if model is None:
                out = voltrix.spmm(*st, M, indices.size, x)
            else:
                kw = {"ft": 64} if (model == 0 and N <= 64) else {}
                if kw:
                    kw.update(stages=20, npw=10)
                voltrix.spmm_kernel(*st, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=x, output=out, model=model, **kw)
            check(f"N={N} {str(dtype)[6:]} model={model}", out.cpu().numpy(), ref, tol if dtype != torch.float32 or model in (1, 2) else 2e-3)
    w = voltrix.edge_weights(*st, torch.from_numpy(indptr), torch.from_numpy(indices), torch.from_numpy(vals))
    x = torch.from_numpy(B).cuda().half()
    got = voltrix.spmm(*st, M, indices.size, x, edge_weights=w)
    check(f"N={N} weighted fp16", got.cpu().numpy(), wtd @ x.float().cpu().numpy(), 3e-3)
torch.cuda.synchronize()
print("SANITIZE_PROBE", "FAILED" if bad else "PASSED")
sys.exit(1 if bad else 0)
