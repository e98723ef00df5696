"""GPU preprocessing is bit-exact with the reference format (SURVEY.md section 8c parity rule i)."""
import ctypes
import hashlib
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import all_case_names, small_case_names

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("name", all_case_names())
def test_csr_preprocess_bit_exact(golden_cases, name):
    import voltrix
    case = golden_cases[name]
    indptr, indices = case["indptr"], case["indices"]
    M = indptr.size - 1
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    assert blk.is_cuda and packed.is_cuda and hind.is_cuda
    assert blk.dtype == torch.int32 and packed.dtype == torch.uint32 and hind.dtype == torch.int32
    npz = case["npz"]
    assert np.array_equal(blk.cpu().numpy(), npz["pointer1"])
    if case["full"]:
        assert np.array_equal(hind.cpu().numpy(), npz["hind"])
        assert np.array_equal(_u32(packed), npz["hspa_packed"])
        assert np.array_equal(packed._vx_plan.block_partition.cpu().numpy(), npz["block_partition"])
    else:
        assert _digest(hind.cpu().numpy()) == str(npz["hind_sha256"])
        assert _digest(_u32(packed)) == str(npz["hspa_packed_sha256"])
    # second element accepts attributes (reference tests/test_spmm.py:55)
    packed.hash_tag = "x"
    # duplicates are detected (they disable the CSR row path)
    plan = packed._vx_plan
    uniq = np.unique(np.stack([np.repeat(np.arange(M), np.diff(indptr)), indices]), axis=1).shape[1] if indices.size else 0
    assert plan.unique_nnz == uniq and plan.has_duplicates == (uniq != indices.size)


@pytest.mark.parametrize("name", small_case_names())
def test_csr_preprocess_deterministic_and_cuda_inputs(golden_cases, name):
    import voltrix
    case = golden_cases[name]
    M = case["indptr"].size - 1
    ip, ix = torch.from_numpy(case["indptr"]).cuda(), torch.from_numpy(case["indices"]).cuda()
    a = voltrix.csr_preprocess(ip, ix, M)
    b = voltrix.csr_preprocess(ip, ix, M)
    for x, y in zip(a, b):
        assert torch.equal(x.view(torch.int32), y.view(torch.int32))


@pytest.mark.parametrize("name", small_case_names() + ["c1_uniform_16384"])
def test_kernel_level_api_matches_reference_outputs(golden_cases, name):
    """preprocess_kernel -> hmat_gen_kernel -> hmat_packed_swizzle_kernel driven by hand with over-allocated
    buffers, as the reference's tests/test_spmm_kernel.py:58-113 does."""
    import voltrix
    case = golden_cases[name]
    indptr, indices = case["indptr"], case["indices"]
    M, E = indptr.size - 1, indices.size
    W = (M + 15) // 16
    # CPU tensors in, CPU tensors out: the reference signature
    bp = torch.zeros(W, dtype=torch.int32)
    e2c = torch.zeros(max(E, 1), dtype=torch.int32)
    e2r = torch.zeros(max(E, 1), dtype=torch.int32)
    p1 = torch.zeros(W + 1, dtype=torch.int32)
    voltrix.preprocess_kernel(edge_list=torch.from_numpy(indices), node_pointer=torch.from_numpy(indptr),
                              block_partition=bp, edge_to_column=e2c, edge_to_row=e2r, pointer1=p1)
    want = oracle.c().preprocess(indptr, indices)
    assert np.array_equal(bp.numpy(), want[0]) and np.array_equal(p1.numpy(), want[3])
    assert np.array_equal(e2c.numpy()[:E], want[1]) and np.array_equal(e2r.numpy()[:E], want[2])

    tcb = int(p1[-1])
    extra = 5
    sentinel = 7.0
    hspa = torch.full(((tcb + extra) * 128,), sentinel, dtype=torch.float32, device="cuda")
    hind = torch.full(((tcb + extra) * 8,), 7, dtype=torch.int32, device="cuda")
    packed = torch.zeros((tcb + extra) * 4, dtype=torch.uint32, device="cuda")
    d = lambda t: t.cuda()
    voltrix.hmat_gen_kernel(node_pointer=d(torch.from_numpy(indptr)), edge_list=d(torch.from_numpy(indices)),
                            block_partition=d(bp), edge_to_column=d(e2c[:E] if E else e2c[:0]),
                            edge_to_row=d(e2r[:E] if E else e2r[:0]), pointer1=d(p1), hspa=hspa, hind=hind)
    voltrix.hmat_packed_swizzle_kernel(block_partition=d(bp), pointer1=d(p1), hspa=hspa, hspa_packed=packed)
    hspa_w, hind_w = oracle.c().hmat(indptr, indices, *want)
    assert np.array_equal(hspa.cpu().numpy()[: tcb * 128], hspa_w)
    assert np.array_equal(hind.cpu().numpy()[: tcb * 8], hind_w)
    # blocks past pointer1[-1] are untouched, like the reference
    assert (hspa.cpu().numpy()[tcb * 128:] == sentinel).all() and (hind.cpu().numpy()[tcb * 8:] == 7).all()
    assert np.array_equal(_u32(packed)[: tcb * 4], oracle.c().pack_swizzle(hspa_w, tcb))


def test_through_the_c_abi_library(golden_cases):
    """Same phases through libvoltrix_b200.so with plain pointers (what a non-Python host would bind)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "voltrix-spmm_b200", "csrc", "libvoltrix_b200.so"))
    case = golden_cases["m1000_sparse"]
    indptr, indices = case["indptr"], case["indices"]
    M, E = indptr.size - 1, indices.size
    W = (M + 15) // 16
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.vx_preprocess_workspace_bytes.restype = ctypes.c_size_t
    lib.vx_preprocess_workspace_bytes.argtypes = [i64, i32]
    nbytes = lib.vx_preprocess_workspace_bytes(E, M)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    ip, ix = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    bp = torch.empty(W, dtype=torch.int32, device="cuda")
    p1 = torch.empty(W + 1, dtype=torch.int32, device="cuda")
    lib.vx_csr_window_sort.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp, ctypes.c_size_t, vp]
    assert lib.vx_csr_window_sort(ip.data_ptr(), ix.data_ptr(), M, E, M, bp.data_ptr(), p1.data_ptr(),
                                  ws.data_ptr(), nbytes, None) == 0
    tcb = int(p1[-1])
    hind = torch.empty(tcb * 8, dtype=torch.int32, device="cuda")
    packed = torch.empty(tcb * 4, dtype=torch.int32, device="cuda")
    uniq = torch.zeros(1, dtype=torch.int64, device="cuda")
    lib.vx_csr_tiles_scatter.argtypes = [i32, i64, i32, vp, i64, vp, vp, vp, vp, ctypes.c_size_t, vp]
    assert lib.vx_csr_tiles_scatter(M, E, M, p1.data_ptr(), tcb, hind.data_ptr(), packed.data_ptr(),
                                    uniq.data_ptr(), ws.data_ptr(), nbytes, None) == 0
    torch.cuda.synchronize()
    npz = case["npz"]
    assert np.array_equal(p1.cpu().numpy(), npz["pointer1"]) and np.array_equal(hind.cpu().numpy(), npz["hind"])
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), npz["hspa_packed"]) and int(uniq) == E
    # workspace too small is reported, not a crash
    assert lib.vx_csr_window_sort(ip.data_ptr(), ix.data_ptr(), M, E, M, bp.data_ptr(), p1.data_ptr(),
                                  ws.data_ptr(), 128, None) == 3

    # SpMM through vx_spmm with nothing but the triple (plan = NULL), all three models
    N = 64
    B = torch.randn(M, N, device="cuda")
    want = oracle.c().spmm_tiles(npz["pointer1"], npz["hspa_packed"], npz["hind"], M, B.cpu().numpy())
    lib.vx_spmm.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp]
    out = torch.empty(M, N, device="cuda")
    assert lib.vx_spmm(p1.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, E, N, B.data_ptr(), 0, out.data_ptr(),
                       2, 16, None, None) == 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=1e-4)
    Bh = B.half()
    want_h = oracle.c().spmm_tiles(npz["pointer1"], npz["hspa_packed"], npz["hind"], M, Bh.float().cpu().numpy())
    for model in (0, 2):
        out.fill_(float("nan"))
        assert lib.vx_spmm(p1.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, E, N, Bh.data_ptr(), 1,
                           out.data_ptr(), model, 16, None, None) == 0
        torch.cuda.synchronize()
        np.testing.assert_allclose(out.cpu().numpy(), want_h, rtol=1e-3, atol=1e-3)
    # model 1 needs the plan's CSR: reported as invalid argument, fp32 has no tcgen05 path: unsupported
    assert lib.vx_spmm(p1.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, E, N, B.data_ptr(), 0, out.data_ptr(),
                       1, 16, None, None) == 1
    assert lib.vx_spmm(p1.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, E, N, B.data_ptr(), 0, out.data_ptr(),
                       0, 16, None, None) == 4
