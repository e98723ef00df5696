"""Transpose SpMM / autograd.  csr_transpose is plain torch and runs on CPU; the backward pass
itself is a GPU test: dX of (A @ X) must equal A^T @ dY, checked against torch.sparse and against finite differences."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch


@pytest.mark.parametrize("shape", [(50, 70), (64, 64), (1, 9), (33, 2)])
def test_csr_transpose_matches_scipy(shape):
    from voltrix.autograd import csr_transpose
    A = sp.random(*shape, density=0.2, format="csr", random_state=np.random.default_rng(0))
    A.data = np.random.default_rng(1).standard_normal(A.nnz).astype(np.float32)
    ip, ix, v = csr_transpose(torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)),
                              shape[1], torch.from_numpy(A.data))
    At = A.T.tocsr(); At.sort_indices()
    assert ip.dtype == torch.int32 and ix.dtype == torch.int32
    assert np.array_equal(ip.numpy(), At.indptr) and np.array_equal(ix.numpy(), At.indices)
    assert np.array_equal(v.numpy(), At.data)
    # num_cols inferred, no values
    ip2, ix2 = csr_transpose(torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)))
    assert np.array_equal(ix2.numpy(), At.indices)        # trailing empty columns cannot be inferred, the entries can


def test_csr_transpose_empty():
    from voltrix.autograd import csr_transpose
    ip, ix = csr_transpose(torch.zeros(5, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), 3)
    assert ip.tolist() == [0, 0, 0, 0] and ix.numel() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_backward_is_the_transpose_spmm(dtype, weighted):
    import voltrix
    M, K, N = 3000, 4100, 64          # rectangular: the transpose has different window geometry
    A = sp.random(M, K, density=0.01, format="csr", random_state=np.random.default_rng(2))
    A.data = (np.random.default_rng(3).integers(1, 9, A.nnz) / 4.0).astype(np.float32) if weighted else np.ones(A.nnz, np.float32)
    ip, ix = torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32))
    adj = voltrix.SparseAdj(ip, ix, M, K, values=torch.from_numpy(A.data) if weighted else None)
    X = torch.randn(K, N, device="cuda").to(dtype).requires_grad_(True)
    Y = adj @ X
    assert Y.shape == (M, N) and Y.dtype == torch.float32
    G = torch.randn(M, N, device="cuda")
    Y.backward(G)
    At = torch.sparse_csr_tensor(ip.cuda(), ix.cuda(), torch.from_numpy(A.data).cuda(), size=(M, K))
    want_Y = At @ X.detach().float()
    want_dX = At.to_dense().T @ G.to(dtype).float()
    tol = 2e-5 if dtype == torch.float32 else 2e-3      # fp16: dX is rounded to the operand dtype
    assert (Y - want_Y).abs().max().item() / want_Y.abs().max().item() <= (2e-5 if dtype == torch.float32 else 1e-4)
    assert X.grad.dtype == dtype
    assert (X.grad.float() - want_dX).abs().max().item() / want_dX.abs().max().item() <= tol
    # no graph when no grad is needed
    with torch.no_grad():
        assert not (adj @ X).requires_grad
