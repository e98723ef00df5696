"""Parity against the REFERENCE'S OWN CUDA KERNELS, compiled unmodified for sm_100a into
oracle/_ref/libvoltrix_ref.so (oracle/ref_harness.cu): hmat_cuda + hmat_packed_swizzle_cuda pin the C oracle's
restatement of the tile format (and therefore the golden hind / hspa_packed vectors), and
voltrix_spmm_forward_cuda (models 0/1/2) is the values reference for the rows it computes."""
import numpy as np
import pytest
import torch

import oracle
from conftest import small_case_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvoltrix_ref.so not present")
    return oracle.ref()


def _ref_tiles(ref, indptr, indices):
    bp, e2c, e2r, p1 = oracle.c().preprocess(indptr, indices)     # == reference preprocess (tests/test_oracle.py)
    M, E, W = indptr.size - 1, indices.size, bp.size
    tcb = int(p1[-1])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t_ip, t_ix = d(indptr), d(indices if E else np.zeros(1, np.int32))
    t_bp, t_e2c, t_e2r, t_p1 = d(bp), d(e2c if E else np.zeros(1, np.int32)), d(e2r if E else np.zeros(1, np.int32)), d(p1)
    hspa = torch.zeros(tcb * 128, device="cuda")
    hind = torch.zeros(tcb * 8, dtype=torch.int32, device="cuda")
    packed = torch.zeros(tcb * 4, dtype=torch.int32, device="cuda")
    assert ref.lib.ref_hmat(t_ip.data_ptr(), t_ix.data_ptr(), t_bp.data_ptr(), t_e2c.data_ptr(), t_e2r.data_ptr(),
                            t_p1.data_ptr(), W, M, E, hspa.data_ptr(), hind.data_ptr()) == 0
    assert ref.lib.ref_hmat_packed_swizzle(W, t_p1.data_ptr(), hspa.data_ptr(), packed.data_ptr()) == 0
    return t_p1, packed, hind, hspa, (bp, e2c, e2r, p1)


@pytest.mark.parametrize("name", small_case_names() + ["c1_uniform_16384"])
def test_tile_format_bit_exact_vs_reference_kernels(ref, golden_cases, name):
    import voltrix
    case = golden_cases[name]
    indptr, indices = case["indptr"], case["indices"]
    M = indptr.size - 1
    r_p1, r_packed, r_hind, r_hspa, pre = _ref_tiles(ref, indptr, indices)
    # (1) the C restatement of the two CUDA kernels is the reference, bit for bit
    c_hspa, c_hind = oracle.c().hmat(indptr, indices, *pre)
    c_packed = oracle.c().pack_swizzle(c_hspa, int(pre[3][-1]))
    assert np.array_equal(r_hspa.cpu().numpy(), c_hspa)
    assert np.array_equal(r_hind.cpu().numpy(), c_hind)
    assert np.array_equal(r_packed.cpu().numpy().view(np.uint32), c_packed)
    # (2) the product's csr_preprocess is the reference, bit for bit
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    assert torch.equal(blk, r_p1) and torch.equal(hind, r_hind)
    assert torch.equal(packed.view(torch.int32), r_packed)


@pytest.mark.parametrize("name", ["m64_dense", "m1000_sparse", "m257_tail1_empty_tail", "c1_uniform_16384"])
@pytest.mark.parametrize("N", [64, 128, 512])
def test_spmm_values_vs_reference_kernel(ref, golden_cases, name, N):
    import voltrix
    from voltrix.utils import calc_diff, relative_error
    case = golden_cases[name]
    indptr, indices = case["indptr"], case["indices"]
    M, E = indptr.size - 1, indices.size
    r_p1, r_packed, r_hind, _, _ = _ref_tiles(ref, indptr, indices)
    feat = torch.randn(M, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(N))
    full = (M // 16) * 16            # the reference computes only whole windows (spmm_kernels.cuh:1514, 2028)
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    for model in (0, 1, 2):
        r_out = torch.zeros(M, N, device="cuda")
        assert ref.lib.ref_spmm(r_p1.data_ptr(), r_packed.data_ptr(), r_hind.data_ptr(), M, E, N, feat.data_ptr(),
                                r_out.data_ptr(), model, torch.cuda.current_stream().cuda_stream) == 0
        torch.cuda.synchronize()
        for dtype in (torch.float32, torch.float16, torch.bfloat16):
            out = voltrix.spmm(blk, packed, hind, M, E, feat.to(dtype))
            a, b = out[:full], r_out[:full]
            # north_star parity bar: <= 1e-2 relative error and a 0.00 % difference rate
            # (reference rounds B to TF32, bf16 keeps 8 mantissa bits: both inside the bar)
            scale = b.abs().max().clamp_min(1e-6)
            assert ((a - b).abs().max() / scale).item() <= 1e-2
            assert f"{abs(calc_diff(a, b).item()) * 100:.2f}" == "0.00"
            if dtype != torch.bfloat16:
                big = b.abs() > 0.05 * scale
                assert relative_error(a[big], b[big]) <= 1e-2
