"""Diagnostic for the tcgen05 path (not a test): single-edge matrices with recognisable B rows, so a wrong
descriptor / swizzle / bitmap-expansion layout shows up as a readable permutation instead of a bare mismatch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
import oracle  # noqa: E402


def run(indptr, indices, M, feat, model=0, stages=16):
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    out = torch.full((M, feat.shape[1]), float("nan"), device="cuda")
    voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=feat.shape[1],
                        input=feat, output=out, model=model, stages=stages)
    torch.cuda.synchronize()
    p1, pk, hi = oracle.c().csr_to_tiles(indptr, indices)
    want = oracle.c().spmm_tiles(p1, pk, hi, M, feat.float().cpu().numpy())
    return out.cpu().numpy(), want


def main():
    M, N = 32, 128
    feat = torch.zeros(M, N)
    for k in range(M):
        feat[k] = k * 128 + torch.arange(N)      # exactly representable in fp16 up to 2048 -> use k < 16
    feat = (feat % 2048).half().cuda()
    ok_all = True
    for (r, c) in [(0, 0), (3, 5), (9, 2), (15, 7), (1, 12), (20, 3)]:
        indptr = np.zeros(M + 1, np.int32)
        indptr[r + 1:] = 1
        indices = np.array([c], np.int32)
        got, want = run(indptr, indices, M, feat)
        ok = np.array_equal(got, want)
        ok_all &= ok
        print(f"edge ({r},{c}): {'OK' if ok else 'MISMATCH'}")
        if not ok:
            nz = np.argwhere(np.nan_to_num(got, nan=-1) != 0)
            print("   nonzero/NaN rows in output:", sorted(set(nz[:, 0].tolist()))[:20])
            rr = nz[0, 0] if len(nz) else r
            print("   got row", rr, got[rr, :16], "...", got[rr, 60:68], "...", got[rr, 120:128])
            print("   want row", r, want[r, :16])
    # two edges same row, different K positions; and a >16-column window (2 K-steps + odd tail)
    rng = np.random.default_rng(0)
    for ncols in (2, 9, 17, 40):
        cols = np.sort(rng.choice(M, size=min(ncols, M), replace=False)).astype(np.int32)
        indptr = np.zeros(M + 1, np.int32)
        indptr[1:] = cols.size
        got, want = run(indptr, cols, M, feat)
        print(f"row0 with {cols.size} cols: max abs err {np.nanmax(np.abs(got - want))}, nan {np.isnan(got).sum()}")
        ok_all &= np.array_equal(got, want)
    # random matrix
    import scipy.sparse as sp
    M2 = 1000
    A = sp.random(M2, M2, density=0.05, format="csr", random_state=rng)
    f2 = torch.randn(M2, 256, device="cuda").half()
    for st in (16, 32):
        got, want = run(A.indptr.astype(np.int32), A.indices.astype(np.int32), M2, f2, stages=st)
        err = np.nanmax(np.abs(got - want)) / np.abs(want).max()
        print(f"random 1000x1000 N=256 stages={st}: scaled err {err:.3e}, nan {np.isnan(got).sum()}")
        ok_all &= err < 1e-4
    print("TC PROBE", "PASS" if ok_all else "FAIL")


if __name__ == "__main__":
    main()
