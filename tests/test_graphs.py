"""Synthetic graph generators (data plumbing for bench.py; run here on the CPU at small sizes)."""
import numpy as np
import torch


def _to_scipy(indptr, indices, M, K=None):
    import scipy.sparse as sp
    return sp.csr_matrix((np.ones(indices.numel(), np.int8), indices.numpy(), indptr.numpy()), shape=(M, K or M))


def test_chung_lu_is_symmetric_coalesced_and_sorted():
    from voltrix.graphs import chung_lu_csr
    M = 3000
    indptr, indices = chung_lu_csr(M, avg_degree=12, max_degree=300, seed=3, device="cpu", target_nnz=36_000)
    A = _to_scipy(indptr, indices, M)
    assert (A != A.T).nnz == 0 and A.diagonal().sum() == 0
    assert abs(A.nnz - 36_000) <= 0.03 * 36_000
    assert all((np.diff(indices.numpy()[indptr[r]:indptr[r + 1]]) > 0).all() for r in range(0, M, 37))
    ip2, ix2 = chung_lu_csr(M, avg_degree=12, max_degree=300, seed=3, device="cpu", target_nnz=36_000)
    assert torch.equal(indptr, ip2) and torch.equal(indices, ix2)          # seeded => every rank builds the same graph


def test_rmat_row_range_shards_tile_the_full_graph():
    from voltrix.graphs import rmat_csr, rmat_row_histogram
    scale, ef = 10, 8
    M = 1 << scale
    ip, ix = rmat_csr(scale, ef, seed=1, device="cpu")
    parts = [rmat_csr(scale, ef, seed=1, device="cpu", row_range=(a, b)) for a, b in ((0, 300), (300, 640), (640, M))]
    assert torch.equal(torch.cat([p[1] for p in parts]), ix)
    offs = np.cumsum([0] + [int(p[0][-1]) for p in parts])
    rebuilt = torch.cat([parts[0][0]] + [p[0][1:] + int(o) for p, o in zip(parts[1:], offs[1:])])
    assert torch.equal(rebuilt, ip)
    hist = rmat_row_histogram(scale, ef, seed=1, device="cpu")
    deg = (ip[1:] - ip[:-1]).long()
    assert int(hist.sum()) == ef * M and bool((hist >= deg).all())            # draw counts bound the coalesced degrees
    assert deg[: M // 2].sum() > deg[M // 2:].sum()                           # a + b = 0.76 of the mass in the top half


def test_named_suite_shapes():
    from voltrix.graphs import named_suite, suite_graph
    names = [n for n, _, _ in named_suite()]
    assert "reddit" in names and len(names) == 12
    ip, ix = suite_graph("ddi", seed=0, device="cpu")
    M, nnz = dict((n, (m, z)) for n, m, z in named_suite())["ddi"]
    assert ip.numel() - 1 == M and abs(ix.numel() - nnz) <= 0.03 * nnz
