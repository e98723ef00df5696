"""On-disk round trip of a preprocessed matrix (host logic on CPU; the GPU test checks that a reloaded matrix multiplies)."""
import numpy as np
import pytest
import torch


def _fake_state():
    from voltrix.spmm.spmm import SpmmPlan
    g = torch.Generator().manual_seed(0)
    blk = torch.tensor([0, 3, 4, 9], dtype=torch.int32)
    packed = torch.randint(0, 2**31 - 1, (9 * 4,), generator=g, dtype=torch.int32).view(torch.uint32)
    hind = torch.randint(0, 48, (9 * 8,), generator=g, dtype=torch.int32)
    plan = SpmmPlan()
    plan.num_nodes, plan.num_edges, plan.total_blocks, plan.unique_nnz = 48, 70, 9, 70
    plan.cap, plan.sparse_ratio, plan.num_items, plan.num_slots, plan.num_fixups, plan.num_sparse_rows = 64, 0.5, 3, 0, 0, 0
    plan.items = torch.tensor([[2, 4, 5, -1], [0, 0, 3, -1], [1, 3, 1, -1]], dtype=torch.int32)
    plan.fixups = torch.zeros((1, 4), dtype=torch.int32)
    plan.sparse_rows = torch.zeros(1, dtype=torch.int32)
    plan.csr_indptr = torch.arange(49, dtype=torch.int32)
    plan.csr_indices = torch.arange(48, dtype=torch.int32)
    plan.block_partition = torch.tensor([3, 1, 5], dtype=torch.int32)
    packed._vx_plan = plan
    packed.hash_tag = "my-graph"
    return blk, packed, hind


def test_round_trip_on_cpu(tmp_path):
    import voltrix
    blk, packed, hind = _fake_state()
    path = str(tmp_path / "g.vxt")
    voltrix.save_preprocessed(path, blk, packed, hind)
    b2, p2, h2 = voltrix.load_preprocessed(path, device="cpu")
    assert torch.equal(b2, blk) and torch.equal(h2, hind)
    assert torch.equal(p2.view(torch.int32), packed.view(torch.int32)) and p2.dtype == torch.uint32
    assert p2.hash_tag == "my-graph"
    a, b = packed._vx_plan, p2._vx_plan
    for k in ("num_nodes", "num_edges", "total_blocks", "cap", "sparse_ratio", "num_items", "num_slots", "num_fixups"):
        assert getattr(a, k) == getattr(b, k)
    for k in ("items", "fixups", "sparse_rows", "csr_indptr", "csr_indices", "block_partition"):
        assert torch.equal(getattr(a, k), getattr(b, k))
    assert b.signature() == a.signature()


def test_rejects_foreign_files(tmp_path):
    import voltrix
    path = str(tmp_path / "x.pt")
    torch.save({"format": "something else", "version": 1}, path)
    with pytest.raises(ValueError):
        voltrix.load_preprocessed(path, device="cpu")


@pytest.mark.gpu
def test_reloaded_matrix_multiplies_identically(tmp_path):
    import voltrix
    from voltrix.graphs import chung_lu_csr
    M, N = 20_000, 128
    indptr, indices = chung_lu_csr(M, avg_degree=25, max_degree=3000, seed=6, device="cuda")
    E = indices.numel()
    st = voltrix.csr_preprocess(indptr, indices, M)
    st[1].hash_tag = "persist-test"
    path = str(tmp_path / "g.vxt")
    voltrix.save_preprocessed(path, *st)
    st2 = voltrix.load_preprocessed(path)
    feat = torch.randn(M, N, device="cuda").half()
    assert torch.equal(voltrix.spmm(*st, M, E, feat), voltrix.spmm(*st2, M, E, feat))
    assert torch.equal(st2[1].view(torch.int32), st[1].view(torch.int32))
