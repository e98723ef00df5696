"""SpMM parity: every kernel path against the oracle on the same inputs; properties at larger sizes."""
import numpy as np
import pytest
import torch

import oracle
from conftest import small_case_names

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.float16, torch.bfloat16]
# parity bars: fp32 CUDA-core paths are exact-fp32 sums (summation order differs from the oracle's: 1e-5 scaled), the
# fp32 tensor-core path (model 3) splits every value into two bf16 terms (|error| <= 2^-18 |x|, well inside 2e-5);
# fp16/bf16 inputs with fp32 accumulation: 1e-2 relative (north_star) -- measured errors are ~1e-6 because
# the oracle is fed the same rounded inputs.
TOL = {torch.float32: 2e-5, torch.float16: 1e-4, torch.bfloat16: 1e-4}


# every tensor-core variant of the autotune space (jit_kernels/spmm.py::SPACE_HALF: 14/7 = three CTAs per SM, 22/11 = two,
# 15/5 = three with 3-stage rings, 42/14 = one) plus the test-only ones (16/4, 40/24 = the two-producer-group geometry):
# (model, K-steps in flight per CTA[, producer warps])
TC_VARIANTS = [(0, 14), (0, 22), (0, 15), (0, 42), (0, 16), (0, 40)]
# the 64-wide feature tile (MMA M = 64) of SPACE_HALF_NARROW, for dense operands of at most 64 columns: (model, stages, npw, ft)
TC_VARIANTS_NARROW = [(0, 20, 10, 64), (0, 21, 7, 64), (0, 33, 11, 64)]


def _scaled_err(got, want):
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-9))


def _run_all_models(voltrix, blk, packed, hind, M, E, feat):
    out = {}
    models = [(1, 32), (2, 32), (3, 24), (3, 12)] if feat.dtype == torch.float32 else TC_VARIANTS + [(1, 32), (2, 32)]
    if feat.dtype != torch.float32 and feat.shape[1] <= 64:
        models = models + TC_VARIANTS_NARROW
    for model, stages, *rest in models:
        o = torch.full((M, feat.shape[1]), float("nan"), device="cuda")
        try:
            voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=E, embedding_dim=feat.shape[1], input=feat,
                                output=o, model=model, stages=stages, npw=rest[0] if rest else None,
                                ft=rest[1] if len(rest) > 1 else None)
        except RuntimeError as e:
            if "invalid argument" in str(e):   # model 1 without CSR (duplicates in the input)
                continue
            if model == 3 and "unsupported" in str(e) and feat.shape[1] % 8 != 0:   # TMA stride rule: N % 8 == 0
                continue
            raise
        out[(model, stages, *rest)] = o.cpu().numpy()
    return out


@pytest.mark.parametrize("name", small_case_names())
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N", [16, 64, 128, 200, 256])
def test_every_path_matches_oracle(golden_cases, name, dtype, N):
    import voltrix
    case = golden_cases[name]
    indptr, indices = case["indptr"], case["indices"]
    M, E = indptr.size - 1, indices.size
    rng = np.random.default_rng(N)
    feat = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    p1, pk, hi = oracle.c().csr_to_tiles(indptr, indices)
    want = oracle.c().spmm_tiles(p1, pk, hi, M, feat.float().cpu().numpy())
    outs = _run_all_models(voltrix, blk, packed, hind, M, E, feat)
    assert outs, "no kernel path ran"
    if not packed._vx_plan.has_duplicates:
        assert any(k[0] == 1 for k in outs), "CSR path must run on coalesced input"
    for key, got in outs.items():
        assert np.isfinite(got).all(), f"{key}: unwritten output rows"
        assert _scaled_err(got, want) <= TOL[dtype], f"model/stages {key}"
    # the public entry point (autotuned) agrees as well and writes the M % 16 tail rows
    got = voltrix.spmm(blk, packed, hind, M, E, feat).cpu().numpy()
    assert _scaled_err(got, want) <= TOL[dtype]


@pytest.mark.parametrize("variant", [(14, None), (22, None), (15, None), (42, None), (16, None), (40, None),
                                     (20, 10, 64), (21, 7, 64), (33, 11, 64)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_sparse_window_routing_and_k_split(dtype, variant):
    """A matrix with one hub window (split along K), many ordinary windows and very sparse windows
    (routed to the CUDA-core row path): all three mechanisms in one SpMM, result equals the oracle."""
    import voltrix
    rng = np.random.default_rng(3)
    ft = variant[2] if len(variant) > 2 else None
    M, N = 4096, (128 if ft is None else 48)      # the 64-wide tile with a width that is not a multiple of 64 either
    rows, cols = [], []
    for r in range(M):
        if r < 16:
            k = 3000                       # hub window: ~all columns
        elif r < 2048:
            k = int(rng.integers(20, 200))
        else:
            k = int(rng.integers(0, 2))    # sparse windows (some rows empty)
        c = rng.choice(M, size=k, replace=False)
        rows.append(np.full(k, r)); cols.append(np.sort(c))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    indptr = np.zeros(M + 1, np.int64); np.add.at(indptr, rows + 1, 1); indptr = np.cumsum(indptr).astype(np.int32)
    indices = cols.astype(np.int32)
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    plan = packed._vx_plan
    assert plan.num_sparse_rows > 0 and plan.num_fixups >= 1 and plan.num_slots >= 2 and plan.num_items > 0
    items = plan.items.cpu().numpy()[: plan.num_items]
    assert (np.diff(items[:, 2]) <= 0).all(), "work list must be LPT (descending block count)"
    feat = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    p1, pk, hi = oracle.c().csr_to_tiles(indptr, indices)
    want = oracle.c().spmm_tiles(p1, pk, hi, M, feat.float().cpu().numpy())
    stages, npw = variant[0], variant[1]
    o = torch.full((M, N), float("nan"), device="cuda")
    voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o,
                        model=0, stages=stages, npw=npw, ft=ft)
    got = o.cpu().numpy()
    assert np.isfinite(got).all()
    assert _scaled_err(got, want) <= 1e-4
    # run-to-run determinism (fixed-order fix-up, no float atomics; which CTA claims which unit does not matter)
    for _ in range(3):
        o2 = torch.empty_like(o)
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o2,
                            model=0, stages=stages, npw=npw, ft=ft)
        assert torch.equal(o, o2)


def test_reference_test_generator_difference_rate():
    """The reference's own check (tests/test_spmm.py:16-33,94): seed 20, sp.random(8192, 8192, 0.1) -- here at
    density 0.01 and N=512 to stay fast -- 'difference rate' vs an fp32 SpMM prints 0.00%."""
    import scipy.sparse as sp
    import voltrix
    from voltrix.utils import calc_diff, relative_error
    np.random.seed(20); torch.manual_seed(20)
    M, N = 8192, 512
    A = sp.random(M, M, density=0.01, format="csr")
    indptr, indices = torch.tensor(A.indptr, dtype=torch.int32), torch.tensor(A.indices, dtype=torch.int32)
    feat = torch.randn(M, N, dtype=torch.float32)
    sparse = torch.sparse_csr_tensor(indptr, indices, values=torch.ones(A.nnz), size=(M, M)).cuda()
    base = (sparse @ feat.cuda())
    blk, packed, hind = voltrix.csr_preprocess(indptr, indices, M)
    packed.hash_tag = "test_20_8192_0.01"
    for dtype in DTYPES:
        f = feat.cuda().to(dtype)
        out = voltrix.spmm(blk, packed, hind, num_nodes=M, num_edges=A.nnz, feat=f)
        assert f"{calc_diff(out, base) * 100:.2f}" in ("0.00", "-0.00")
        # north_star bar: 1e-2 relative error of an fp32 SpMM of the same (rounded) inputs.  relative_error is the
        # reference's MEAN ELEMENTWISE metric (utils.py:21-35); against the un-rounded fp32 B it is dominated by
        # outputs that cancel to ~0 (randn B), so bf16's 2^-9 input rounding alone reads 1.5e-2 there.
        assert relative_error(out, sparse @ f.float()) <= 1e-2
        if dtype != torch.bfloat16:
            assert relative_error(out, base) <= 1e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_properties_at_scale(dtype):
    """Size-independent properties on a 200k-row power-law graph the CPU oracle would take minutes on:
    linearity in B, agreement between independent kernel paths, row-degree identity with B = ones."""
    import voltrix
    from voltrix.graphs import chung_lu_csr
    M, N = 200_000, 128
    indptr, indices = chung_lu_csr(M, avg_degree=40, max_degree=4000, seed=1, device="cuda")
    E = indices.numel()
    blk, packed, hind = voltrix.csr_preprocess(indptr, indices, M)
    ones = torch.ones(M, N, device="cuda", dtype=dtype)
    deg = (indptr[1:] - indptr[:-1]).float()
    out1 = voltrix.spmm(blk, packed, hind, M, E, ones)
    assert torch.equal(out1, deg[:, None].expand(M, N)), "A @ ones must equal the row degrees exactly"
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(M, N, device="cuda", generator=g).to(dtype)
    Y = torch.randn(M, N, device="cuda", generator=g).to(dtype)
    fx = voltrix.spmm(blk, packed, hind, M, E, X)
    fy = voltrix.spmm(blk, packed, hind, M, E, Y)
    fxy = voltrix.spmm(blk, packed, hind, M, E, (X.float() + Y.float()).to(dtype))
    scale = fxy.abs().max().item()
    # fp16: X+Y is rounded to fp16 once more.  fp32: the default precision class is the reference's (TF32-like: the
    # operand may travel as one fp16 term, 2^-12 relative per value); VOLTRIX_FP32_MODE=split / exact tighten it (below)
    tol = 2e-2 if dtype == torch.float16 else 2e-3
    assert (fxy - (fx + fy)).abs().max().item() / scale < tol
    # independent paths agree (tensor-core vs CSR rows vs tile rows)
    ref = None
    for model in ((1, 0, 2) if dtype != torch.float32 else (1, 2, 3)):
        o = torch.empty(M, N, device="cuda")
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=E, embedding_dim=N, input=X, output=o, model=model)
        if ref is None:
            ref = o
        else:
            assert (o - ref).abs().max().item() / scale < (2e-5 if model == 3 else 1e-5)
    # against cuSPARSE fp32 (the reference's comparison, tests/test_spmm.py:75-85)
    sparse = torch.sparse_csr_tensor(indptr, indices, torch.ones(E, device="cuda"), size=(M, M))
    base = sparse @ X.float()
    assert (fx - base).abs().max().item() / scale < (1e-4 if dtype == torch.float16 else 1e-3)
    if dtype == torch.float32:      # the tighter precision classes, selected by policy (not by which candidate was fastest)
        import os
        for mode, bar in (("split", 2e-5), ("exact", 1e-5)):
            os.environ["VOLTRIX_FP32_MODE"] = mode
            try:
                got = voltrix.spmm(blk, packed, hind, M, E, X)
            finally:
                os.environ.pop("VOLTRIX_FP32_MODE", None)
            assert (got - base).abs().max().item() / scale < bar, mode


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N", [32, 64, 128, 512])
@pytest.mark.parametrize("degree", [1.7, 3.0, 5.0])
def test_low_degree_rows_use_group_per_row_kernel(dtype, N, degree):
    """Mean degree 2-5 (the Yeast / DD end of the C3 suite): the CSR path switches to one lane group per row
    (vx_spmm_csr_subwarp_rows_kernel; one warp per row at N = 512).  Result equals the oracle's CSR SpMM -- bit for bit in
    fp32, the sum being in CSR order -- including a few long rows, rows without non-zeros (written as 0) and a row count
    that is not a multiple of the rows per warp."""
    import scipy.sparse as sp
    import voltrix
    M = 5003
    A = sp.random(M, M, density=degree / M, format="lil", random_state=np.random.default_rng(7))
    rng = np.random.default_rng(N)
    for r in (1, 2500, M - 1):                     # long rows among the short ones
        for c in rng.choice(M, size=301, replace=False):
            A[r, c] = 1.0
    A = A.tocsr(); A.sort_indices()
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    assert (np.diff(indptr) == 0).any() and indices.size / M < 6.0
    feat = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    want = oracle.c().spmm_csr(indptr, indices, feat.float().cpu().numpy(), 0, M, assume_coalesced=True)
    o = torch.full((M, N), float("nan"), device="cuda")
    voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o,
                        model=1)
    got = o.cpu().numpy()
    assert np.isfinite(got).all()
    assert _scaled_err(got, want) <= TOL[dtype]
    if dtype == torch.float32:
        short = np.diff(indptr) <= 8               # the oracle sums in the same order; long rows differ by FMA contraction at most
        assert np.array_equal(got[short], want[short])
    # the same rows through the public call: sparse windows go to the CSR kernel by row list, the rest to tcgen05
    auto = voltrix.spmm(blk, packed, hind, M, indices.size, feat)
    assert _scaled_err(auto.cpu().numpy(), want) <= max(TOL[dtype], 2e-3 if dtype == torch.float32 else 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("N", [32, 128, 512])
def test_short_rows_at_scale_take_several_rows_per_warp(dtype, N):
    """Above 2^17 rows of mean degree < 8 the CSR launchers give every warp (or lane group) FOUR rows: row ids and CSR
    bounds of all four are loaded first, then the rows are walked (spmm_cuda_core.cuh, RPW).  Same result as the oracle, bit
    for bit in fp32; row count not a multiple of 4 or 32; empty rows written as zeros; also through the row list of the
    sparse windows (public call) and with per-entry values."""
    import scipy.sparse as sp
    import voltrix
    M = (1 << 17) + 24581
    rng = np.random.default_rng(N)
    deg = rng.integers(0, 5, size=M); deg[rng.integers(0, M, 40)] = 70            # mostly 0-4, a few longer rows
    pattern = sp.coo_matrix((np.ones(int(deg.sum()), np.float32), (np.repeat(np.arange(M), deg), rng.integers(0, M, int(deg.sum())))),
                            shape=(M, M)).tocsr()
    pattern.sum_duplicates(); pattern.sort_indices()
    indptr, indices = pattern.indptr.astype(np.int32), pattern.indices.astype(np.int32)
    deg = np.diff(indptr)
    assert (deg == 0).any() and M % 4 and M % 32
    feat = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    want = oracle.c().spmm_csr(indptr, indices, feat.float().cpu().numpy(), 0, M, assume_coalesced=True)
    o = torch.full((M, N), float("nan"), device="cuda")
    voltrix.spmm_kernel(*st, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o, model=1)
    got = o.cpu().numpy()
    assert np.isfinite(got).all() and _scaled_err(got, want) <= TOL[dtype]
    if dtype == torch.float32:
        assert np.array_equal(got[deg <= 8], want[deg <= 8])
    voltrix.reschedule(*st, 4.0, 0)                  # route (nearly) every window to the CUDA-core rows: the row-list launch
    assert st[1]._vx_plan.num_sparse_rows >= (1 << 17)                                # ... is RPW too
    auto = voltrix.spmm(*st, M, indices.size, feat)
    assert _scaled_err(auto.cpu().numpy(), want) <= max(TOL[dtype], 2e-3 if dtype == torch.float32 else 0)
    vals = rng.uniform(0.5, 1.5, indices.size).astype(np.float32)
    A = sp.csr_matrix((vals, indices, indptr), shape=(M, M))
    w = voltrix.spmm_weighted(torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda(), torch.from_numpy(vals).cuda(), feat)
    assert _scaled_err(w.cpu().numpy(), A @ feat.float().cpu().numpy()) <= 1e-4


def test_fp32_tensor_core_path_precision_and_range():
    """Model 3 keeps 16 mantissa bits and the full fp32 exponent range: values far outside fp16's range and a
    mantissa pattern that TF32 (10 bits) would round away both survive."""
    import scipy.sparse as sp
    import voltrix
    M, N = 2048, 128
    A = sp.random(M, M, density=0.02, format="csr", random_state=np.random.default_rng(11))
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    rng = np.random.default_rng(5)
    B = (rng.standard_normal((M, N)) * 10.0 ** rng.integers(-20, 20, size=(M, 1))).astype(np.float32)   # per-row scales
    B[:, 0] = 1.0 + 2.0 ** -14                                                                           # needs > 10 bits
    feat = torch.from_numpy(B).cuda()
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    want = oracle.c().spmm_csr(indptr, indices, B, 0, M, assume_coalesced=True)
    o = torch.full((M, N), float("nan"), device="cuda")
    voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o,
                        model=3)
    got = o.cpu().numpy()
    assert np.isfinite(got).all()
    deg = np.diff(indptr)
    col0 = got[:, 0]
    assert np.abs(col0 - deg * (1.0 + 2.0 ** -14)).max() <= 1e-6 * max(deg.max(), 1), "lost low mantissa bits"
    # rows mix magnitudes 1e-20 .. 1e20: compare per output element relative to the largest addend of that element
    Ad = sp.csr_matrix((np.ones(indices.size, np.float32), indices, indptr), shape=(M, M))
    bound = (Ad @ np.abs(B)).astype(np.float64) + 1e-300
    assert (np.abs(got.astype(np.float64) - want) / bound).max() <= 2e-5


def test_host_streamed_pipeline_matches_direct_calls():
    """voltrix.HostStreamedSpMM: pinned-host operands, copies and kernels of consecutive steps overlapped on three
    streams; every step's result equals the direct call's bit for bit, in submission order."""
    import voltrix
    from voltrix.graphs import chung_lu_csr
    M, N = 20_000, 128
    indptr, indices = chung_lu_csr(M, avg_degree=30, max_degree=2000, seed=2, device="cuda")
    E = indices.numel()
    st = voltrix.csr_preprocess(indptr, indices, M)
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(M, N, generator=g).half().pin_memory() for _ in range(5)]
    outs = [torch.empty(M, N).pin_memory() for _ in range(5)]
    pipe = voltrix.HostStreamedSpMM(*st, M, E, N, dtype=torch.float16)
    for f, o in zip(feats, outs):
        pipe.submit(f, o)
    pipe.wait()
    for f, o in zip(feats, outs):
        want = voltrix.spmm(*st, M, E, f.cuda()).cpu()
        assert torch.equal(o, want)


def _epilogue_case():
    """One hub window (K-split -> fix-up pass), ordinary windows (tensor-core epilogue), sparse windows (CSR rows), M % 16 != 0."""
    rng = np.random.default_rng(9)
    M = 3001
    rows, cols = [], []
    for r in range(M):
        k = 2500 if r < 16 else (int(rng.integers(20, 120)) if r < 1500 else int(rng.integers(0, 2)))
        rows.append(np.full(k, r)); cols.append(np.sort(rng.choice(M, size=k, replace=False)))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    indptr = np.zeros(M + 1, np.int64); np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr).astype(np.int32), cols.astype(np.int32), M


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_epilogue_every_model(dtype):
    """act(row_scale * (A @ B) + bias) applied inside each kernel that writes C equals the unfused result post-processed
    on the host side -- for every model, including K-split windows (fix-up pass) and CUDA-core sparse rows."""
    import voltrix
    indptr, indices, M = _epilogue_case()
    N = 128
    rng = np.random.default_rng(1)
    feat = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    scale = torch.from_numpy(rng.uniform(0.1, 2.0, M).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).cuda()
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    plan = packed._vx_plan
    assert plan.num_fixups >= 1 and plan.num_sparse_rows > 0
    models = [(1, 32), (2, 32), (3, 24), (3, 12), (4, 12)] if dtype == torch.float32 else \
        [(0, 14), (0, 22), (0, 15), (0, 42), (0, 40), (1, 32), (2, 32)]
    for model, stages in models:
        plain = torch.empty(M, N, device="cuda")
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat,
                            output=plain, model=model, stages=stages)
        for rs, b, relu in ((scale, None, False), (None, bias, False), (None, None, True), (scale, bias, True)):
            o = torch.full((M, N), float("nan"), device="cuda")
            voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat,
                                output=o, model=model, stages=stages, row_scale=rs, bias=b, relu=relu)
            want = plain * (rs[:, None] if rs is not None else 1.0) + (b[None, :] if b is not None else 0.0)
            if relu:
                want = want.clamp_min(0.0)
            assert torch.isfinite(o).all()
            err = (o - want).abs().max().item() / max(want.abs().max().item(), 1e-9)
            assert err <= 1e-6, (model, stages, rs is not None, b is not None, relu, err)   # one fma of difference


def test_gcn_layer_matches_dense_formula():
    """voltrix.spmm_gcn = relu(D^-1/2 A D^-1/2 X + b) against torch.sparse on the normalised matrix (fp32)."""
    import voltrix
    from voltrix.graphs import chung_lu_csr
    M, N = 30_000, 64
    indptr, indices = chung_lu_csr(M, avg_degree=20, max_degree=1500, seed=4, device="cuda")
    E = indices.numel()
    st = voltrix.csr_preprocess(indptr, indices, M)
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(M, N, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    dinv = voltrix.gcn_norm(indptr)
    got = voltrix.spmm_gcn(*st, M, E, X, dinv, bias=b, relu=True)
    rows = torch.repeat_interleave(torch.arange(M, device="cuda"), (indptr[1:] - indptr[:-1]).long())
    vals = dinv[rows] * dinv[indices.long()]
    A = torch.sparse_csr_tensor(indptr, indices, vals, size=(M, M))
    want = torch.relu(A @ X + b)
    assert (got - want).abs().max().item() / want.abs().max().item() <= 5e-5


def test_c_abi_plan_with_epilogue():
    """vx_spmm from libvoltrix_b200.so with a vx_plan_t built by hand (work list + fused epilogue): the struct layout of
    include/voltrix_b200.h is what the library reads."""
    import ctypes
    import os
    import voltrix
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = ctypes.CDLL(os.path.join(ROOT, "voltrix-spmm_b200", "csrc", "libvoltrix_b200.so"))

    class Plan(ctypes.Structure):
        _fields_ = [("items", ctypes.c_void_p), ("num_items", ctypes.c_int32), ("fixups", ctypes.c_void_p),
                    ("num_fixups", ctypes.c_int32), ("scratch", ctypes.c_void_p), ("csr_indptr", ctypes.c_void_p),
                    ("csr_indices", ctypes.c_void_p), ("sparse_rows", ctypes.c_void_p), ("num_sparse_rows", ctypes.c_int32),
                    ("input_rows", ctypes.c_int64), ("split_ws", ctypes.c_void_p), ("row_scale", ctypes.c_void_p),
                    ("bias", ctypes.c_void_p), ("relu", ctypes.c_int32), ("ticket", ctypes.c_void_p),
                    ("value_tiles", ctypes.c_void_p), ("csr_values", ctypes.c_void_p), ("sparse_mean_degree", ctypes.c_float)]

    indptr, indices, M = _epilogue_case()
    N = 128
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    p = packed._vx_plan
    feat = torch.randn(M, N, device="cuda").half()
    scale = torch.rand(M, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    scratch = p.scratch(N)
    plan = Plan(p.items.data_ptr(), p.num_items, p.fixups.data_ptr(), p.num_fixups, scratch.data_ptr() if scratch is not None else None,
                p.csr_indptr.data_ptr(), p.csr_indices.data_ptr(), p.sparse_rows.data_ptr(), p.num_sparse_rows, M, None,
                scale.data_ptr(), bias.data_ptr(), 1, None, None, None, 0.0)
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    lib.vx_spmm.restype = ctypes.c_int
    lib.vx_spmm.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, vp, i32, i32, ctypes.POINTER(Plan), vp]
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.vx_spmm(blk.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, indices.size, N, feat.data_ptr(), 1,
                     out.data_ptr(), 0, 14, ctypes.byref(plan), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    want = torch.relu(voltrix.spmm(blk, packed, hind, M, indices.size, feat) * scale[:, None] + bias[None, :])
    assert torch.isfinite(out).all()
    assert (out - want).abs().max().item() / want.abs().max().item() <= 1e-6
    # ABI v5: with a ticket counter the persistent CTAs claim their units dynamically; same bits out
    ticket = torch.full((4,), 12345, dtype=torch.int32, device="cuda")   # vx_spmm zeroes it on the stream
    plan.ticket = ticket.data_ptr()
    out2 = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.vx_spmm(blk.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, indices.size, N, feat.data_ptr(), 1,
                     out2.data_ptr(), 0, 22, ctypes.byref(plan), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    assert int(ticket[0]) >= plan.num_items    # every unit beyond the first grid-full was claimed through the counter


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_cuda_graph_capture_and_replay(dtype):
    """The whole SpMM (tensor-core kernel + fix-up + CUDA-core rows, or split + tensor-core for fp32) is capturable in a CUDA
    graph once the variant is tuned: nothing on the launch path allocates or synchronises.  Replays see new operand values."""
    import voltrix
    indptr, indices, M = _epilogue_case()
    N = 128
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    E = indices.size
    feat = torch.randn(M, N, device="cuda").to(dtype)
    out = torch.empty(M, N, device="cuda")
    voltrix.spmm(*st, M, E, feat, out=out)          # tune / load outside the capture
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            voltrix.spmm(*st, M, E, feat, out=out)
    torch.cuda.current_stream().wait_stream(side)
    for seed in (1, 2):
        feat.copy_(torch.randn(M, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed)).to(dtype))
        out.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        want = voltrix.spmm(*st, M, E, feat)
        assert torch.equal(out, want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("density,N", [(3.0 / 4000, 32), (0.02, 128), (0.02, 512)])
def test_weighted_csr_spmm(dtype, density, N):
    """voltrix.spmm_weighted (general CSR values, CUDA-core rows, both the warp-per-row and the group-per-row kernel)
    against scipy on the same rounded operand; with and without the fused epilogue."""
    import scipy.sparse as sp
    import voltrix
    M = 4000
    A = sp.random(M, M, density=density, format="csr", random_state=np.random.default_rng(3), dtype=np.float32)
    A.data = np.random.default_rng(4).standard_normal(A.nnz).astype(np.float32)
    feat = torch.from_numpy(np.random.default_rng(5).standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    ip, ix, vals = (torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)),
                    torch.from_numpy(A.data))
    want = A @ feat.float().cpu().numpy()
    got = voltrix.spmm_weighted(ip.cuda(), ix.cuda(), vals.cuda(), feat).cpu().numpy()
    assert _scaled_err(got, want) <= 2e-5
    scale = torch.rand(M, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    got2 = voltrix.spmm_weighted(ip, ix, vals, feat, row_scale=scale, bias=bias, relu=True).cpu().numpy()
    want2 = np.maximum(want * scale.cpu().numpy()[:, None] + bias.cpu().numpy()[None, :], 0.0)
    assert _scaled_err(got2, want2) <= 2e-5


def test_reference_timing_hook_agrees_with_cuda_events():
    """The reference's timing call -- GPU_bench(fn, kernel_name="spmm") (bench/bm_voltrix.py:36, utils.py:232-303): profiler
    time of the kernels whose name contains "spmm", L2 flushed per iteration -- must give the per-call kernel time: every
    kernel an SpMM launches carries "spmm" in its name, and the figure agrees with CUDA events around the same calls."""
    import voltrix
    from voltrix.graphs import chung_lu_csr
    from voltrix.utils import GPU_bench
    M, N = 100_000, 128
    indptr, indices = chung_lu_csr(M, avg_degree=60, max_degree=30_000, seed=6, device="cuda")   # hub windows: K-split + fix-up
    E = indices.numel()
    st = voltrix.csr_preprocess(indptr, indices, M)
    assert st[1]._vx_plan.num_fixups >= 1
    feat = torch.rand(M, N, device="cuda").half()
    fn = lambda: voltrix.spmm(*st, M, E, feat)   # noqa: E731
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    events_ms = float(np.median(ts))
    hook_ms = GPU_bench(fn, iters=10, warmup=10, kernel_name="spmm")
    assert 0.6 * events_ms <= hook_ms <= 1.1 * events_ms, (hook_ms, events_ms)   # kernel time <= event time (launch gaps)
