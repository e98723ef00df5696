"""Synthetic graph generators for the benchmark configurations.

The reference benchmarks on downloaded datasets loaded through TC-GNN's ``TCGNN_dataset``
(bench/graph_gen.py:47-49); there is no network here, so the named shapes are synthesised:

* ``uniform_csr``   -- C1: ``scipy.sparse.random`` uniform pattern (the reference tests' generator).
* ``chung_lu_csr``  -- C2 / C4: power-law expected-degree graph (Chung-Lu), symmetric, randomly labelled,
                       coalesced, sorted columns.  ``reddit_shaped`` / ``products_shaped`` pick the sizes.
* ``rmat_csr``      -- C5: R-MAT (a, b, c, d) edge stream, coalesced; optionally only a row range, so each
                       rank of a multi-GPU run can build just its shard.

All generators run on whatever torch device they are given (GPU for the real sizes -- a numpy Chung-Lu of
120 M edges took ~300 s during the survey) and return int32 (indptr, indices) tensors on that device.
Generation is data plumbing, not part of the timed path.
"""
import math
from typing import Optional, Tuple

import numpy as np
import torch


def uniform_csr(M: int, nnz: int, seed: int = 0, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """C1: scipy.sparse.random(M, M, density=nnz/M^2, random_state=default_rng(seed)); values ignored."""
    import scipy.sparse as sp
    A = sp.random(M, M, density=nnz / float(M) ** 2, format="csr", random_state=np.random.default_rng(seed))
    return (torch.from_numpy(A.indptr.astype(np.int32)).to(device),
            torch.from_numpy(A.indices.astype(np.int32)).to(device))


def _csr_from_keys(keys: torch.Tensor, M: int, col_bits: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """keys = (row << col_bits) | col, int64, any order, duplicates allowed -> coalesced CSR with sorted columns."""
    keys = torch.unique(keys)          # sorts and coalesces
    rows = keys >> col_bits
    cols = (keys & ((1 << col_bits) - 1)).to(torch.int32)
    del keys
    indptr = torch.searchsorted(rows, torch.arange(M + 1, device=rows.device, dtype=rows.dtype)).to(torch.int32)
    return indptr, cols


def power_law_weights(M: int, avg_degree: float, max_degree: float, device="cpu") -> torch.Tensor:
    """Expected degrees w_i = max_degree * (i + 1)^-alpha with alpha fitted so that mean(w) = avg_degree."""
    ranks = torch.arange(1, M + 1, dtype=torch.float64, device=device)
    lo, hi = 0.0, 4.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        mean = (max_degree * ranks.pow(-mid)).mean().item()
        if mean > avg_degree:
            lo = mid
        else:
            hi = mid
    return max_degree * ranks.pow(-0.5 * (lo + hi))


def chung_lu_csr(M: int, avg_degree: float, max_degree: float, seed: int = 0, device="cpu",
                 target_nnz: Optional[int] = None, chunk: int = 1 << 26) -> Tuple[torch.Tensor, torch.Tensor]:
    """Symmetric Chung-Lu graph: each undirected edge picks both endpoints with probability ~ w_i.

    Node labels are a random permutation of the degree ranks (hubs are NOT adjacent -- the hard case for
    column sharing).  ``target_nnz`` (directed entries after coalescing) is met within
    ~1 % by one calibration pass on the duplicate rate.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    w = power_law_weights(M, avg_degree, max_degree, dev)
    cdf = torch.cumsum(w / w.sum(), 0).to(torch.float64)
    label = torch.randperm(M, generator=g, device=dev)
    col_bits = max(1, int(M - 1).bit_length())
    want = int(target_nnz) if target_nnz is not None else int(round(avg_degree * M))

    def draw(n_undirected: int) -> torch.Tensor:
        parts = []
        left = n_undirected
        while left > 0:
            n = min(left, chunk)
            u = torch.searchsorted(cdf, torch.rand(n, generator=g, device=dev, dtype=torch.float64)).clamp_(max=M - 1)
            v = torch.searchsorted(cdf, torch.rand(n, generator=g, device=dev, dtype=torch.float64)).clamp_(max=M - 1)
            u, v = label[u], label[v]
            keep = u != v
            u, v = u[keep], v[keep]
            parts.append(torch.unique(torch.cat([(u << col_bits) | v, (v << col_bits) | u])))
            left -= n
        return torch.unique(torch.cat(parts)) if len(parts) > 1 else parts[0]

    n0 = want // 2
    keys = draw(n0)
    got = keys.numel()
    if target_nnz is not None and abs(got - want) > 0.01 * want:
        # duplicate / self-loop losses grow with density: rescale the draw count once
        del keys
        keys = draw(int(n0 * (want / got) * (1 + 0.5 * (want / got - 1))))
    return _csr_from_keys(keys, M, col_bits)


def reddit_shaped(seed: int = 0, device="cuda", scale: float = 1.0):
    """C2: M = 232 965, ~114.6 M nnz, mean degree ~492, max degree ~21 657.  ``scale`` < 1 shrinks
    the node count (same degree shape) for quick runs."""
    M = max(64, int(232_965 * scale))
    avg = 114_615_892 / 232_965
    return chung_lu_csr(M, avg_degree=min(avg, M / 4), max_degree=min(21_657, M / 2), seed=seed, device=device,
                        target_nnz=int(min(avg, M / 4) * M))


def products_shaped(seed: int = 0, device="cuda", scale: float = 1.0):
    """C4: M = 2 449 029, ~123.7 M nnz (mean degree ~50.5, max degree ~17 481 as in ogbn-products)."""
    M = max(64, int(2_449_029 * scale))
    avg = 123_718_280 / 2_449_029
    return chung_lu_csr(M, avg_degree=avg, max_degree=min(17_481, M / 2), seed=seed, device=device,
                        target_nnz=int(avg * M))


def rmat_edge_chunk(scale: int, n: int, probs, gen: torch.Generator, device) -> Tuple[torch.Tensor, torch.Tensor]:
    a, b, c, _ = probs
    rows = torch.zeros(n, dtype=torch.int64, device=device)
    cols = torch.zeros(n, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(n, generator=gen, device=device)
        rb = r >= (a + b)                          # quadrants c, d -> row bit
        cb = ((r >= a) & (r < a + b)) | (r >= a + b + c)   # quadrants b, d -> col bit
        rows = (rows << 1) | rb
        cols = (cols << 1) | cb
    return rows, cols


def rmat_csr(scale: int, edge_factor: int = 32, probs=(0.57, 0.19, 0.19, 0.05), seed: int = 0, device="cuda",
             row_range: Optional[Tuple[int, int]] = None, chunk: int = 1 << 26):
    """C5: R-MAT with 2^scale rows and edge_factor * 2^scale edge draws, coalesced.

    With ``row_range=(r0, r1)`` only rows in [r0, r1) are kept (row ids stay global in the stream but the
    returned indptr covers r1 - r0 rows); every rank replays the same seeded stream, so shards are consistent.
    Returns (indptr, indices, row_degree_hint) where row_degree_hint is None unless row_range is None.
    """
    dev = torch.device(device)
    M = 1 << scale
    total = edge_factor * M
    g = torch.Generator(device=dev).manual_seed(seed)
    r0, r1 = row_range if row_range is not None else (0, M)
    parts = []
    done = 0
    while done < total:
        n = min(chunk, total - done)
        rows, cols = rmat_edge_chunk(scale, n, probs, g, dev)
        if row_range is not None:
            keep = (rows >= r0) & (rows < r1)
            rows, cols = rows[keep] - r0, cols[keep]
        parts.append(torch.unique((rows << scale) | cols))
        del rows, cols
        done += n
        if len(parts) >= 8:   # bound the number of live chunks
            parts = [torch.unique(torch.cat(parts))]
    keys = torch.unique(torch.cat(parts)) if len(parts) > 1 else parts[0]
    del parts
    return _csr_from_keys(keys, r1 - r0, scale)


def rmat_row_histogram(scale: int, edge_factor: int = 32, probs=(0.57, 0.19, 0.19, 0.05), seed: int = 0,
                       device="cuda", chunk: int = 1 << 26) -> torch.Tensor:
    """Row draw counts of the same seeded R-MAT stream (pre-coalescing) -- enough to nnz-balance row ranges
    across ranks without materialising the whole graph on one GPU."""
    dev = torch.device(device)
    M = 1 << scale
    total = edge_factor * M
    g = torch.Generator(device=dev).manual_seed(seed)
    hist = torch.zeros(M, dtype=torch.int64, device=dev)
    done = 0
    while done < total:
        n = min(chunk, total - done)
        rows, _ = rmat_edge_chunk(scale, n, probs, g, dev)
        hist += torch.bincount(rows, minlength=M)
        done += n
    return hist


def named_suite():
    """C3: (name, M, nnz) of the GNN / SuiteSparse graphs in the reference's plot (bench/plot.py:8), matched by
    size only -- the edge structure is a Chung-Lu stand-in (max degree = min(M/8, 40 * mean))."""
    return [
        ("ddi", 4_267, 2_135_822), ("ppi", 56_944, 818_716), ("protein", 43_471, 162_088), ("DD", 334_925, 1_686_092),
        ("amazon0505", 410_236, 4_878_874), ("amazon0601", 403_394, 5_478_357), ("com-amazon", 334_863, 1_851_744),
        ("web-BerkStan", 685_230, 7_600_595), ("Yeast", 1_710_902, 3_636_546), ("YeastH", 3_138_114, 6_487_230),
        ("FraudYelp-RSR", 45_954, 6_805_486), ("reddit", 232_965, 114_615_892),
    ]


def suite_graph(name: str, seed: int = 0, device="cuda"):
    for n, M, nnz in named_suite():
        if n == name:
            avg = nnz / M
            return chung_lu_csr(M, avg_degree=avg, max_degree=max(avg * 2, min(M / 8, avg * 40)), seed=seed,
                                device=device, target_nnz=nnz)
    raise KeyError(name)


def planted_partition_csr(M: int, community: int, p_in: float, p_out: float, seed: int = 0, device="cpu",
                          shuffle: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Symmetric stochastic block model with equal communities of ``community`` nodes (edge probability ``p_in`` inside,
    ``p_out`` across), node labels shuffled when ``shuffle`` -- the structure real GNN graphs have and the Chung-Lu
    stand-ins lack; used to exercise voltrix.reorder."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    C = (M + community - 1) // community
    col_bits = max(1, int(M - 1).bit_length())
    parts = []
    # inside: per community a dense Bernoulli block (upper triangle), across: a global sparse sample
    n_in = int(C * community * (community - 1) / 2 * p_in * 1.05) + 16
    c = torch.randint(0, C, (n_in,), generator=g, device=dev)
    u = c * community + torch.randint(0, community, (n_in,), generator=g, device=dev)
    v = c * community + torch.randint(0, community, (n_in,), generator=g, device=dev)
    parts.append((u, v))
    n_out = int(M * (M - community) / 2 * p_out) + 16
    parts.append((torch.randint(0, M, (n_out,), generator=g, device=dev), torch.randint(0, M, (n_out,), generator=g, device=dev)))
    u = torch.cat([p[0] for p in parts]); v = torch.cat([p[1] for p in parts])
    keep = (u != v) & (u < M) & (v < M)
    u, v = u[keep], v[keep]
    if shuffle:
        label = torch.randperm(M, generator=g, device=dev)
        u, v = label[u], label[v]
    return _csr_from_keys(torch.cat([(u << col_bits) | v, (v << col_bits) | u]), M, col_bits)
