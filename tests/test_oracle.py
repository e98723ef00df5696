"""Pins the oracle: C restatement == numpy restatement == committed golden vectors (which come from the
reference's own voltrix::preprocess) == the reference itself when oracle/_ref is available."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import all_case_names, small_case_names


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", all_case_names())
def test_c_oracle_matches_golden(golden_cases, name):
    case = golden_cases[name]
    bp, e2c, e2r, p1 = oracle.c().preprocess(case["indptr"], case["indices"])
    hspa, hind = oracle.c().hmat(case["indptr"], case["indices"], bp, e2c, e2r, p1)
    packed = oracle.c().pack_swizzle(hspa, int(p1[-1]))
    got = dict(block_partition=bp, edge_to_column=e2c, edge_to_row=e2r, pointer1=p1, hind=hind, hspa_packed=packed)
    npz = case["npz"]
    if case["full"]:
        assert np.array_equal(npz["indptr"], case["indptr"]) and np.array_equal(npz["indices"], case["indices"])
        for k, v in got.items():
            assert np.array_equal(npz[k], v), k
    else:
        assert str(npz["indptr_sha256"]) == _digest(case["indptr"]), "generator drifted (numpy/scipy version?)"
        assert str(npz["indices_sha256"]) == _digest(case["indices"])
        for k, v in got.items():
            assert str(npz[k + "_sha256"]) == _digest(v), k


@pytest.mark.parametrize("name", small_case_names())
def test_numpy_restatement_matches_c(golden_cases, name):
    case = golden_cases[name]
    c = oracle.c().preprocess(case["indptr"], case["indices"])
    n = oracle.np_preprocess(case["indptr"], case["indices"])
    for a, b in zip(c, n):
        assert np.array_equal(a, b)
    hspa, hind, packed = oracle.np_tiles(case["indptr"], case["indices"], *n)
    hspa_c, hind_c = oracle.c().hmat(case["indptr"], case["indices"], *c)
    assert np.array_equal(hspa, hspa_c) and np.array_equal(hind, hind_c)
    assert np.array_equal(packed, oracle.c().pack_swizzle(hspa_c, int(c[3][-1])))


@pytest.mark.parametrize("name", small_case_names())
def test_reference_preprocess_matches_c(golden_cases, name, capfd):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built and reference tree not mounted")
    case = golden_cases[name]
    r = oracle.ref().preprocess(case["indptr"], case["indices"])
    capfd.readouterr()  # the reference printf's its TC block count
    c = oracle.c().preprocess(case["indptr"], case["indices"])
    for a, b in zip(r, c):
        assert np.array_equal(a, b)


def test_empty_window_owns_one_block(golden_cases):
    npz = golden_cases["m48_all_empty"]["npz"]
    assert npz["block_partition"].tolist() == [1, 1, 1] and npz["pointer1"].tolist() == [0, 1, 2, 3]
    assert not npz["hspa_packed"].any() and not npz["hind"].any()


def test_swizzle_bit_order():
    """bit beta of word idx <-> tile[(beta>>2) + 8*(idx&1)][(beta&3) + 4*(idx>>1)] (bmat_kernels.cuh:180-184)."""
    for r in range(16):
        for cc in range(8):
            hspa = np.zeros(128, np.float32)
            hspa[r * 8 + cc] = 1.0
            packed = oracle.c().pack_swizzle(hspa, 1)
            idx, beta = (r >> 3) + 2 * (cc >> 2), ((r & 7) << 2) + (cc & 3)
            want = np.zeros(4, np.uint32)
            want[idx] = np.uint32(1) << np.uint32(beta)
            assert np.array_equal(packed, want)


@pytest.mark.parametrize("name", small_case_names())
def test_spmm_tiles_matches_scipy(golden_cases, name):
    case = golden_cases[name]
    M = case["indptr"].size - 1
    rng = np.random.default_rng(1)
    B = rng.standard_normal((M, 48)).astype(np.float32)
    p1, packed, hind = oracle.c().csr_to_tiles(case["indptr"], case["indices"])
    got = oracle.c().spmm_tiles(p1, packed, hind, M, B)
    want = oracle.np_spmm_binary(case["indptr"], case["indices"], B)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-4)
    # the reference kernel leaves the M % 16 tail rows untouched (spmm_kernels.cuh:1514, 2028)
    ref_like = oracle.c().spmm_tiles(p1, packed, hind, M, B, all_windows=False)
    full = (M // 16) * 16
    np.testing.assert_array_equal(ref_like[:full], got[:full])
    assert not ref_like[full:].any()
    # TF32 rounding of B (cvt.rna, spmm_kernels.cuh:1671-1672) stays within the 1e-2 contract
    tf = oracle.c().spmm_tiles(p1, packed, hind, M, B, round_tf32=True)
    denom = np.maximum(np.abs(want), 1.0)
    assert (np.abs(tf - want) / denom).max() < 1e-2


def test_spmm_csr_port_binary_semantics(golden_cases):
    case = golden_cases["tiny_37_unsorted_dups"]   # unsorted + duplicates: sort rows first, duplicates count once
    indptr, indices = case["indptr"], case["indices"].copy()
    for r in range(indptr.size - 1):
        indices[indptr[r]:indptr[r + 1]].sort()
    B = np.random.default_rng(2).standard_normal((37, 16)).astype(np.float32)
    got = oracle.c().spmm_csr(indptr, indices, B)
    np.testing.assert_allclose(got, oracle.np_spmm_binary(indptr, indices, B), rtol=1e-5, atol=1e-5)
