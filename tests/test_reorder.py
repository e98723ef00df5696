"""voltrix.reorder: min-hash LSH relabelling (host logic, runs on the CPU with torch ops)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle


def _graph(M=2048, community=64, seed=0):
    from voltrix.graphs import planted_partition_csr
    return planted_partition_csr(M, community, p_in=0.3, p_out=0.0005, seed=seed)


def test_tc_block_count_matches_oracle_tiles():
    from voltrix import reorder
    indptr, indices = _graph()
    p1, _, _ = oracle.c().csr_to_tiles(indptr.numpy(), indices.numpy())
    assert reorder.tc_block_count(indptr, indices) == int(p1[-1])
    # an edgeless window still counts one block (reference rule)
    ip = torch.tensor([0] * 17 + [1] * 16, dtype=torch.int32); ix = torch.tensor([5], dtype=torch.int32)
    p1, _, _ = oracle.c().csr_to_tiles(ip.numpy(), ix.numpy())
    assert reorder.tc_block_count(ip, ix) == int(p1[-1]) == 2


def test_lsh_reorder_is_a_permutation_and_deterministic():
    from voltrix import reorder
    indptr, indices = _graph()
    M = indptr.numel() - 1
    perm = reorder.lsh_reorder(indptr, indices)
    assert torch.equal(torch.sort(perm).values, torch.arange(M))
    assert torch.equal(perm, reorder.lsh_reorder(indptr, indices))


def test_permuted_product_equals_original():
    from voltrix import reorder
    indptr, indices = _graph(M=1000, community=50)
    M = indptr.numel() - 1
    perm = reorder.lsh_reorder(indptr, indices)
    ip2, ix2 = reorder.permute_graph(indptr, indices, perm)
    assert ip2[-1] == indptr[-1]
    A = sp.csr_matrix((np.ones(indices.numel(), np.float32), indices.numpy(), indptr.numpy()), shape=(M, M))
    A2 = sp.csr_matrix((np.ones(ix2.numel(), np.float32), ix2.numpy(), ip2.numpy()), shape=(M, M))
    assert (np.diff(ip2.numpy()) == np.diff(indptr.numpy())[perm.numpy()]).all()
    assert all((np.diff(ix2.numpy()[ip2[r]:ip2[r + 1]]) > 0).all() for r in range(0, M, 97))   # sorted columns
    B = torch.randn(M, 8)
    want = A @ B.numpy()
    got = reorder.unpermute_rows(torch.from_numpy(A2 @ reorder.permute_rows(B, perm).numpy()), perm).numpy()
    assert np.allclose(got, want, atol=1e-4)


def test_lsh_reorder_recovers_planted_communities():
    """Shuffled labels scatter every community over all windows; sorting by min-hash signature brings rows with common
    neighbours back together: ~2x fewer TC blocks (= gathered B rows); 4324 -> 2044 on this graph, the unshuffled graph
    needs 1322."""
    from voltrix import reorder
    indptr, indices = _graph()
    before = reorder.tc_block_count(indptr, indices)
    perm = reorder.lsh_reorder(indptr, indices, num_hashes=2)
    after = reorder.tc_block_count(*reorder.permute_graph(indptr, indices, perm))
    ideal = reorder.tc_block_count(*__import__("voltrix").graphs.planted_partition_csr(2048, 64, 0.3, 0.0005, seed=0, shuffle=False))
    assert after * 1.8 <= before and ideal <= after, (before, after, ideal)
    # degree ordering alone does not find the structure
    after_deg = reorder.tc_block_count(*reorder.permute_graph(indptr, indices, reorder.degree_reorder(indptr)))
    assert after < after_deg
