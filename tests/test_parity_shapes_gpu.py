"""Parity of the SHIPPED path on the BASELINE shapes: the autotuned ``voltrix.spmm`` (whatever variant the tuner picks --
14/7 or 22/11 on the large graphs) and every tensor-core variant of the tune space, against the CPU oracle on scaled instances of
C2 (Reddit-shaped), C4 (products-shaped) and C5 (a non-square R-MAT row shard with hub windows that get K-split).

Same comparison as the reference's own test (tests/test_spmm.py:75-96: ``calc_diff`` "difference rate" against an fp32 SpMM,
expected 0.00 %), asserted here, plus the north_star bar (1e-2 relative) and the tighter scaled-error bar of the other GPU
tests.  Also: dense widths that are not a multiple of the vector width, rectangular A, two streams on one matrix, Inf in
the fp32 tensor-core path, N = 512 across four feature tiles.
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

HALF_VARIANTS = [(14, 7), (22, 11), (15, 5), (42, 14)]   # the tensor-core points of SPACE_HALF (3 / 2 / 3 / 1 CTAs per SM)


def _scaled_err(got, want):
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-9))


def _shape(name):
    """(indptr, indices, rows of A, rows of B, N) on the GPU, >= 2 M non-zeros each."""
    from voltrix import graphs
    if name == "c2_reddit_scaled":          # M = 23 296, ~11 M nnz, mean degree ~490 (dense windows, L2-resident B)
        ip, ix = graphs.reddit_shaped(seed=0, device="cuda", scale=0.1)
        return ip, ix, ip.numel() - 1, ip.numel() - 1, 128
    if name == "c4_products_scaled":        # M = 122 451, ~6.2 M nnz, mean degree ~50, N = 256 (two feature tiles)
        ip, ix = graphs.products_shaped(seed=0, device="cuda", scale=0.05)
        return ip, ix, ip.numel() - 1, ip.numel() - 1, 256
    if name == "c5_rmat_shard":             # rows [0, 16384) of a scale-19 R-MAT: 16 384 x 524 288, ~3.6 M nnz, hub rows
        ip, ix = graphs.rmat_csr(19, 32, seed=0, device="cuda", row_range=(0, 16384))    # of ~40 k non-zeros first
        return ip, ix, 16384, 1 << 19, 256
    raise KeyError(name)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name", ["c2_reddit_scaled", "c4_products_scaled", "c5_rmat_shard"])
def test_autotuned_and_every_tc_variant_match_oracle(name, dtype):
    import voltrix
    from voltrix.utils import calc_diff, relative_error
    indptr, indices, M, K, N = _shape(name)
    E = indices.numel()
    assert E >= 2_000_000
    blk, packed, hind = voltrix.csr_preprocess(indptr, indices, M, num_cols=K)
    plan = packed._vx_plan
    if name == "c5_rmat_shard":
        assert plan.num_fixups >= 1, "the hub windows of the R-MAT shard must be K-split"
    g = torch.Generator(device="cuda").manual_seed(1)
    feat = torch.rand(K, N, device="cuda", generator=g).to(dtype)      # uniform[0,1) like bench/graph_gen.py:66
    want = oracle.c().spmm_csr(indptr.cpu().numpy(), indices.cpu().numpy(), feat.float().cpu().numpy(), 0, M,
                               assume_coalesced=True)
    want_t = torch.from_numpy(want).cuda()

    def check(out, what):
        assert torch.isfinite(out).all(), what
        assert _scaled_err(out.cpu().numpy(), want) <= 1e-4, what
        assert f"{calc_diff(out, want_t) * 100:.2f}" in ("0.00", "-0.00"), what     # the reference's printed check
        assert relative_error(out, want_t) <= 1e-2, what                            # north_star bar

    out = voltrix.spmm(blk, packed, hind, M, E, feat)          # the public, autotuned entry point
    check(out, f"autotuned {voltrix.jit_tuner.tuned_keys}")
    assert torch.equal(out, voltrix.spmm(blk, packed, hind, M, E, feat)), "run-to-run determinism"
    for stages, npw in HALF_VARIANTS:
        o = torch.full((M, N), float("nan"), device="cuda")
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=E, embedding_dim=N, input=feat, output=o,
                            model=0, stages=stages, npw=npw)
        check(o, f"model 0 variant {stages}/{npw}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N", [1, 7, 100, 130])
def test_dense_width_not_a_multiple_of_the_vector_width(dtype, N):
    """N = 100 in fp16 has no 16-byte row slices and no TMA-legal row stride: the autotuned entry point must still answer
    (scalar-access CUDA-core rows), and so must the explicit CUDA-core models and the weighted CSR kernel."""
    import scipy.sparse as sp
    import voltrix
    M = 3000
    A = sp.random(M, M, density=0.01, format="csr", random_state=np.random.default_rng(N))
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    feat = torch.from_numpy(np.random.default_rng(1).standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    want = oracle.c().spmm_csr(indptr, indices, feat.float().cpu().numpy(), 0, M, assume_coalesced=True)
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    got = voltrix.spmm(blk, packed, hind, M, indices.size, feat)
    assert torch.isfinite(got).all() and _scaled_err(got.cpu().numpy(), want) <= 2e-5
    for model in (1, 2):
        o = torch.full((M, N), float("nan"), device="cuda")
        voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o,
                            model=model)
        assert torch.isfinite(o).all() and _scaled_err(o.cpu().numpy(), want) <= 2e-5, model
    vals = torch.ones(indices.size, device="cuda")
    got_w = voltrix.spmm_weighted(torch.from_numpy(indptr), torch.from_numpy(indices), vals, feat)
    assert _scaled_err(got_w.cpu().numpy(), want) <= 2e-5
    if N % 8 != 0:   # the tensor-core model itself says "unsupported" instead of mis-running
        with pytest.raises(RuntimeError, match="unsupported"):
            voltrix.spmm_kernel(blk, packed, hind, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat.half(),
                                output=torch.empty(M, N, device="cuda"), model=0)


def test_rectangular_matrix_through_the_three_argument_signature():
    """More columns than rows, num_cols not given (the reference's signature has no such argument and its std::map
    compaction takes any column id): tiles bit-exact against the oracle, SpMM against the oracle, and an explicit num_cols
    that is too small is refused instead of corrupting memory."""
    import scipy.sparse as sp
    import voltrix
    M, K, N = 500, 70_000, 64
    A = sp.random(M, K, density=0.002, format="csr", random_state=np.random.default_rng(3))
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    assert indices.max() >= 65_536 > M
    blk, packed, hind = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    p1, pk, hi = oracle.c().csr_to_tiles(indptr, indices)
    assert np.array_equal(blk.cpu().numpy(), p1) and np.array_equal(hind.cpu().numpy(), hi)
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), pk)
    feat = torch.randn(K, N, device="cuda").half()
    want = oracle.c().spmm_csr(indptr, indices, feat.float().cpu().numpy(), 0, M, assume_coalesced=True)
    got = voltrix.spmm(blk, packed, hind, M, indices.size, feat)
    assert got.shape == (M, N) and _scaled_err(got.cpu().numpy(), want) <= 1e-4
    with pytest.raises(ValueError, match="out of range"):
        voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M, num_cols=M)
    bad = indices.copy(); bad[0] = -1
    with pytest.raises(ValueError, match="negative"):
        voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(bad), M)


def test_two_streams_on_the_same_matrix():
    """spmm() from two CUDA streams on one preprocessed matrix (K-split scratch, ticket counter and the fp32 split
    workspace are per stream): both results equal the single-stream result bit for bit."""
    import voltrix
    from test_spmm_gpu import _epilogue_case
    indptr, indices, M = _epilogue_case()          # hub window (K-split), ordinary and sparse windows
    N = 256
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    assert st[1]._vx_plan.num_fixups >= 1
    for dtype in (torch.float16, torch.float32):
        feats = [torch.randn(M, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(s)).to(dtype)
                 for s in (1, 2)]
        want = [voltrix.spmm(*st, M, indices.size, f) for f in feats]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        for rep in range(20):
            outs = []
            for s, f in zip(streams, feats):
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    outs.append(voltrix.spmm(*st, M, indices.size, f))
            torch.cuda.synchronize()
            assert torch.equal(outs[0], want[0]) and torch.equal(outs[1], want[1]), (dtype, rep)


def test_fp32_tensor_core_path_keeps_inf_and_nan():
    """Model 3 splits x = hi + lo in bf16: for Inf (or a finite value that rounds to bf16 Inf) lo must be 0, not
    Inf - Inf = NaN.  A row WITH an edge to an Inf row of B gets Inf, as on the exact-fp32 path; a NaN stays a NaN.
    (Rows of the same window WITHOUT that edge see 0 x Inf inside the MMA tile -- a property of every dense-tile SpMM,
    the reference's mma.sync kernels included -- so they are not part of the claim.)"""
    import scipy.sparse as sp
    import voltrix
    M, N = 512, 64
    A = sp.random(M, M, density=0.03, format="csr", random_state=np.random.default_rng(2))
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    B = np.random.default_rng(3).random((M, N)).astype(np.float32)      # non-negative: no Inf - Inf in a row sum
    B[7, :] = np.inf
    B[9, 0] = 3.4e38          # finite in fp32, rounds to bf16 Inf
    B[11, 1] = np.nan
    feat = torch.from_numpy(B).cuda()
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    split = torch.empty(M, N, device="cuda")
    voltrix.spmm_kernel(*st, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=split, model=3)
    split = split.cpu().numpy()
    dense = A.toarray() != 0
    has7, has9, has11 = dense[:, 7], dense[:, 9], dense[:, 11]
    assert has7.any() and has9.any() and has11.any()
    assert np.isposinf(split[has7][:, 2:]).all(), "Inf x 1 must stay +Inf (lo term must not be Inf - Inf)"
    # feature 0 of B row 9 overflows bf16: rows with that edge, in windows that gather neither the Inf row nor the NaN
    win_has = lambda col: np.repeat(dense[:, col].reshape(-1, 16).any(1), 16)      # noqa: E731
    rows9 = has9 & ~win_has(7) & ~win_has(11)
    assert rows9.any()
    assert (np.isposinf(split[rows9, 0]) | (split[rows9, 0] > 3e38)).all()
    assert np.isnan(split[has11, 1]).all()


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_four_feature_tiles_with_k_split_and_sparse_rows(dtype):
    """N = 512: units are claimed feature-tile-major (all items of columns 0..127, then 128..255, ...); K-split partial
    tiles, the fix-up pass and the CUDA-core sparse rows all see every tile."""
    import voltrix
    from test_spmm_gpu import _epilogue_case
    indptr, indices, M = _epilogue_case()
    N = 512
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    feat = torch.from_numpy(np.random.default_rng(0).standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    want = oracle.c().spmm_csr(indptr, indices, feat.float().cpu().numpy(), 0, M, assume_coalesced=True)
    variants = [(0, 14), (0, 22), (0, 15), (0, 42), (0, 40)] if dtype == torch.float16 else [(3, 24), (3, 12), (4, 12)]
    for model, stages in variants:
        o = torch.full((M, N), float("nan"), device="cuda")
        voltrix.spmm_kernel(*st, num_nodes=M, num_edges=indices.size, embedding_dim=N, input=feat, output=o, model=model,
                            stages=stages)
        assert torch.isfinite(o).all()
        tol = 1e-4 if dtype == torch.float16 else (5e-4 if model == 4 else 2e-5)     # model 4 carries fp32 as one fp16 term
        assert _scaled_err(o.cpu().numpy(), want) <= tol, (model, stages)


def _weighted_case(dtype_exact=True, seed=5):
    """The epilogue case (hub window -> K-split, ordinary windows, sparse windows -> CUDA-core rows, M % 16 != 0) with a value
    per stored entry.  Values are multiples of 1/8 in [-4, 4]: exactly representable in fp16 AND bf16, so rounding them into
    the 16-bit value tiles is lossless and the tensor-core result can be held to the same bar as the binary path."""
    from test_spmm_gpu import _epilogue_case
    indptr, indices, M = _epilogue_case()
    rng = np.random.default_rng(seed)
    vals = (rng.integers(-32, 33, size=indices.size) / 8.0).astype(np.float32)
    vals[vals == 0] = 0.125
    return indptr, indices, vals, M


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N", [32, 64, 128, 256])
def test_weighted_tensor_core_path_matches_scipy(dtype, N):
    """A with per-edge values on the tcgen05 kernel (value tiles bulk-copied straight into the A^T operand stage) -- every
    tensor-core variant, the weighted CUDA-core model and the autotuned entry point against scipy on the same operand;
    with the fused epilogue; K-split windows and sparse windows included."""
    import scipy.sparse as sp
    import voltrix
    indptr, indices, vals, M = _weighted_case()
    E = indices.size
    A = sp.csr_matrix((vals, indices, indptr), shape=(M, M))
    feat = torch.from_numpy(np.random.default_rng(1).standard_normal((M, N)).astype(np.float32)).cuda().to(dtype)
    want = A @ feat.float().cpu().numpy()
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    plan = st[1]._vx_plan
    assert plan.num_fixups >= 1 and plan.num_sparse_rows > 0
    w = voltrix.edge_weights(*st, torch.from_numpy(indptr), torch.from_numpy(indices), torch.from_numpy(vals))
    got = voltrix.spmm(*st, M, E, feat, edge_weights=w)
    assert torch.isfinite(got).all() and _scaled_err(got.cpu().numpy(), want) <= 1e-4
    variants = [(0, 16, None, None), (0, 14, None, None), (0, 22, None, None), (0, 15, None, None), (0, 42, None, None),
                (1, 32, None, None)]
    if N <= 64:      # the 64-wide feature tile takes the value tiles as well
        variants += [(0, 20, 10, 64), (0, 21, 7, 64), (0, 33, 11, 64)]
    for model, stages, npw, ft in variants:
        o = torch.full((M, N), float("nan"), device="cuda")
        voltrix.spmm_kernel(*st, num_nodes=M, num_edges=E, embedding_dim=N, input=feat, output=o, model=model,
                            stages=stages, npw=npw, edge_weights=w, ft=ft)
        assert torch.isfinite(o).all(), (model, stages)
        assert _scaled_err(o.cpu().numpy(), want) <= 1e-4, (model, stages, npw)
    # fused epilogue on top of the weighted product
    scale = torch.rand(M, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    o = voltrix.spmm(*st, M, E, feat, edge_weights=w, row_scale=scale, bias=bias, relu=True)
    want2 = np.maximum(want * scale.cpu().numpy()[:, None] + bias.cpu().numpy()[None, :], 0.0)
    assert _scaled_err(o.cpu().numpy(), want2) <= 1e-4
    # the binary product of the same triple is untouched by the weights
    plain = voltrix.spmm(*st, M, E, feat)
    want_plain = sp.csr_matrix((np.ones(E, np.float32), indices, indptr), shape=(M, M)) @ feat.float().cpu().numpy()
    assert _scaled_err(plain.cpu().numpy(), want_plain) <= 1e-4


def test_weighted_fp32_operand_and_general_values():
    """fp32 dense operand: the weighted product runs on the exact-fp32 CUDA-core rows.  General (not 16-bit representable)
    values on the tensor cores are rounded to the operand's format: within fp16's 2^-11 of scipy."""
    import scipy.sparse as sp
    import voltrix
    indptr, indices, _, M = _weighted_case()
    E = indices.size
    vals = np.random.default_rng(2).standard_normal(E).astype(np.float32)
    A = sp.csr_matrix((vals, indices, indptr), shape=(M, M))
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    w = voltrix.edge_weights(*st, torch.from_numpy(indptr), torch.from_numpy(indices), torch.from_numpy(vals))
    f32 = torch.from_numpy(np.random.default_rng(3).standard_normal((M, 128)).astype(np.float32)).cuda()
    got = voltrix.spmm(*st, M, E, f32, edge_weights=w)
    assert _scaled_err(got.cpu().numpy(), A @ f32.cpu().numpy()) <= 2e-5
    f16 = f32.half()
    o = torch.empty(M, 128, device="cuda")
    voltrix.spmm_kernel(*st, num_nodes=M, num_edges=E, embedding_dim=128, input=f16, output=o, model=0, stages=14,
                        edge_weights=w)
    assert _scaled_err(o.cpu().numpy(), A @ f16.float().cpu().numpy()) <= 1e-3


def test_edge_weights_refuses_foreign_triples_and_duplicates():
    import voltrix
    indptr, indices, vals, M = _weighted_case()
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    # a CSR matrix that is NOT the one the triple was built from: row 16 gets a column its window (rows 16..31) never had
    window_cols = set(indices[indptr[16]:indptr[32]].tolist())
    foreign = next(c for c in range(M) if c not in window_cols)
    other = indices.copy()
    row = other[indptr[16]:indptr[17]].copy(); row[0] = foreign; other[indptr[16]:indptr[17]] = np.sort(row)
    w = voltrix.edge_weights(*st, torch.from_numpy(indptr), torch.from_numpy(other), torch.from_numpy(vals))
    with pytest.raises(ValueError, match="no slot"):
        w.tiles(torch.float16)
    dup_ptr = np.array([0, 3], np.int32); dup_idx = np.array([1, 1, 2], np.int32)
    st2 = voltrix.csr_preprocess(torch.from_numpy(dup_ptr), torch.from_numpy(dup_idx), 1, num_cols=4)
    with pytest.raises(ValueError, match="more than once"):
        voltrix.edge_weights(*st2, torch.from_numpy(dup_ptr), torch.from_numpy(dup_idx), torch.ones(3))


def test_fp32_single_fp16_term_path_and_its_range_fallback():
    """Model 4: an fp32 operand inside fp16's normal range runs as ONE fp16 term (11 significant bits -- the reference rounds
    to TF32's 10, spmm_kernels.cuh:1631-1678); one value outside the range flips the device-side flag and the same call takes
    the two-term bf16 pipeline (16 bits).  Both within the north_star bar; the in-range result equals the fp16 kernel fed the
    rounded operand bit for bit; the out-of-range result equals model 3's."""
    import scipy.sparse as sp
    import voltrix
    from test_spmm_gpu import _epilogue_case
    indptr, indices, M = _epilogue_case()
    N, E = 128, indices.size
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    B = np.random.default_rng(4).standard_normal((M, N)).astype(np.float32)
    B[np.abs(B) < 1e-3] = 0.25                           # keep every value inside fp16's normal range
    feat = torch.from_numpy(B).cuda()
    want = oracle.c().spmm_csr(indptr, indices, B, 0, M, assume_coalesced=True, acc64=True)

    def run(model, x):
        o = torch.full((M, N), float("nan"), device="cuda")
        voltrix.spmm_kernel(*st, num_nodes=M, num_edges=E, embedding_dim=N, input=x, output=o, model=model)
        return o

    got = run(4, feat)
    assert torch.isfinite(got).all()
    assert _scaled_err(got.cpu().numpy(), want) <= 5e-4              # 2^-12 relative per operand value
    assert voltrix.utils.relative_error(got, torch.from_numpy(want).cuda()) <= 1e-2
    as_f16 = torch.full((M, N), float("nan"), device="cuda")
    voltrix.spmm_kernel(*st, num_nodes=M, num_edges=E, embedding_dim=N, input=feat.half(), output=as_f16, model=0, stages=14)
    tc_rows = torch.ones(M, dtype=torch.bool, device="cuda")
    tc_rows[st[1]._vx_plan.sparse_rows[: st[1]._vx_plan.num_sparse_rows].long()] = False     # sparse rows stay exact fp32
    assert torch.equal(got[tc_rows], as_f16[tc_rows])
    # the carrier is scaled by a power of two first: an operand that is tiny (or huge) as a whole is still one fp16 term
    for k in (2.0 ** -40, 2.0 ** 30):
        got_k = run(4, feat * k)
        assert _scaled_err(got_k.cpu().numpy(), want * np.float32(k)) <= 5e-4
        assert torch.equal(got_k[tc_rows], (got * k)[tc_rows])             # power-of-two scaling is exact end to end
    # a dynamic range fp16 cannot span (2^-14 .. 2^15 after scaling): same call, other pipeline
    feat2 = feat.clone(); feat2[17, 3] = 1.0e9
    B2 = feat2.cpu().numpy()
    want2 = oracle.c().spmm_csr(indptr, indices, B2, 0, M, assume_coalesced=True, acc64=True)
    got2 = run(4, feat2)
    assert torch.isfinite(got2).all() and _scaled_err(got2.cpu().numpy(), want2) <= 2e-5
    assert torch.equal(got2, run(3, feat2))
    # and back: the flag is re-evaluated on every call
    assert torch.equal(run(4, feat), got)
    # stray values close to zero (any large Gaussian operand has some) do not re-route the call: the gate counts groups
    # with a value below fp16's normal range, it does not look at the minimum
    feat3 = feat.clone(); feat3[11, 5] = 1.0e-12; feat3[900, 77] = -3.0e-20
    got3 = run(4, feat3)
    want3 = oracle.c().spmm_csr(indptr, indices, feat3.cpu().numpy(), 0, M, assume_coalesced=True, acc64=True)
    assert _scaled_err(got3.cpu().numpy(), want3) <= 5e-4
    as_f16_3 = torch.full((M, N), float("nan"), device="cuda")
    voltrix.spmm_kernel(*st, num_nodes=M, num_edges=E, embedding_dim=N, input=feat3.half(), output=as_f16_3, model=0, stages=14)
    assert torch.equal(got3[tc_rows], as_f16_3[tc_rows])            # still the one-term pipeline
    assert not torch.equal(got3, run(3, feat3))
    # rows of B far below the rest are structure, not stray values (8 groups of 128 values here; 3 are tolerated)
    feat4 = feat.clone(); feat4[40:48] *= 1.0e-11
    assert torch.equal(run(4, feat4), run(3, feat4))


@pytest.mark.parametrize("N", [32, 64, 48])
def test_fp32_single_fp16_term_narrow_widths(N):
    """Model 4 at N <= 64 runs its fp16 term on the 64-wide feature tile (tcgen05.mma M = 64); same bar as N = 128, the
    range fallback still lands on model 3's result, and a width that is not a multiple of 16 takes the masked epilogue."""
    import voltrix
    from test_spmm_gpu import _epilogue_case
    indptr, indices, M = _epilogue_case()
    E = indices.size
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    B = np.random.default_rng(40 + N).standard_normal((M, N)).astype(np.float32)
    B[np.abs(B) < 1e-3] = 0.25
    feat = torch.from_numpy(B).cuda()
    want = oracle.c().spmm_csr(indptr, indices, B, 0, M, assume_coalesced=True, acc64=True)

    def run(model, x):
        o = torch.full((M, N), float("nan"), device="cuda")
        voltrix.spmm_kernel(*st, num_nodes=M, num_edges=E, embedding_dim=N, input=x, output=o, model=model)
        return o

    got = run(4, feat)
    assert torch.isfinite(got).all()
    assert _scaled_err(got.cpu().numpy(), want) <= 5e-4
    feat2 = feat.clone(); feat2[5, 1] = 3.0e9
    want2 = oracle.c().spmm_csr(indptr, indices, feat2.cpu().numpy(), 0, M, assume_coalesced=True, acc64=True)
    got2 = run(4, feat2)
    assert _scaled_err(got2.cpu().numpy(), want2) <= 2e-5
    assert torch.equal(got2, run(3, feat2))
    assert torch.equal(run(4, feat), got)


def test_reschedule_and_tune_routing():
    """The routing rule (which windows leave the tensor cores for the CUDA-core rows) can be changed on a finished triple
    (voltrix.reschedule: schedule kernels only) and picked by measurement (voltrix.tune_routing).  Every
    rule computes the same product; prepared launches and K-split scratch of the old work list are dropped."""
    import voltrix
    from test_spmm_gpu import _epilogue_case
    indptr, indices, M = _epilogue_case()
    N, E = 128, indices.size
    st = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M)
    plan = st[1]._vx_plan
    triple_before = [t.clone() for t in st]
    feat = torch.from_numpy(np.random.default_rng(8).standard_normal((M, N)).astype(np.float32)).cuda().half()
    want = oracle.c().spmm_csr(indptr, indices, feat.float().cpu().numpy(), 0, M, assume_coalesced=True, acc64=True)
    base = voltrix.spmm(*st, M, E, feat)
    assert plan.route_is_default and plan.num_sparse_rows > 0
    default_sparse = plan.num_sparse_rows
    voltrix.reschedule(*st, 0.0, 0)                       # everything on the tensor cores
    assert plan.num_sparse_rows == 0 and not plan.route_is_default
    all_tc = voltrix.spmm(*st, M, E, feat)
    voltrix.reschedule(*st, 4.0, 64)                      # far more windows on the CUDA-core rows
    assert plan.num_sparse_rows > default_sparse and not plan.route_is_default
    mostly_rows = voltrix.spmm(*st, M, E, feat)
    for got in (base, all_tc, mostly_rows):
        assert _scaled_err(got.cpu().numpy(), want) <= 1e-4
    best, timings = voltrix.tune_routing(*st, M, E, feat, iters=3)
    assert best in timings and len(timings) >= 2 and timings[best] == min(timings.values())
    assert (plan.sparse_ratio, plan.small_blocks) == best
    assert _scaled_err(voltrix.spmm(*st, M, E, feat).cpu().numpy(), want) <= 1e-4
    for a, b in zip(st, triple_before):                   # the reference-format triple never changes
        assert torch.equal(a, b)
    # a triple without the CSR arrays has nothing to route: no-op
    st2 = voltrix.csr_preprocess(torch.from_numpy(indptr), torch.from_numpy(indices), M, keep_csr=False)
    voltrix.reschedule(*st2, 1.0, 32)
    assert st2[1]._vx_plan.num_sparse_rows == 0
    assert voltrix.tune_routing(*st2, M, E, feat) == (None, {})
