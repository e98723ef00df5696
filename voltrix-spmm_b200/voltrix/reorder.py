"""Row/node reordering for tile density.

The reference is benchmarked on graphs relabelled offline by DTC-SpMM's ``TCA_reorder.py`` (min-hash LSH over the
neighbour sets, datasketch + cugraph on the CPU; ``bench/bench_all.py:23,120-149``, ``bench/graph_gen.py:42-45`` load the
resulting ``*.reorder.npz``).  Rows that share neighbours land in the same 16-row window, their columns compact into
fewer 16x8 TC blocks, and every gathered B row serves several window rows -- the one lever that lowers the gather bytes
that bound the SpMM.

This module does the same job on the GPU with torch ops only (device-agnostic: the tests run it on the CPU):

* ``minhash_signatures``  -- K min-hash values per row: ``min over the row's columns of (a_k * col + b_k) mod p``.
* ``lsh_reorder``         -- permutation that sorts nodes by their signature (rows with the same min-hash neighbour
                             become adjacent; ties broken by the next hash, then by node id => deterministic).
* ``cluster_reorder``     -- window-aware agglomerative clustering: rounds of "every cluster proposes to its most similar
                             candidate (estimated Jaccard from min-hash signatures), mutual proposals merge, clusters close
                             at 16 rows" -- the parallel counterpart of TCA_reorder.py's greedy priority queue.  Slower than
                             ``lsh_reorder`` (a few hundred ms on a 10^5-row graph) and denser tiles.
* ``degree_reorder``      -- permutation by descending degree (cheap baseline; groups the hubs).
* ``permute_graph``       -- relabel a square adjacency matrix, rows AND columns: ``A' = P A P^T``.  Then
                             ``A' (P B) = P (A B)``: permute the rows of B with ``perm`` going in and read row ``i`` of the
                             result as node ``perm[i]`` (helpers ``permute_rows`` / ``unpermute_rows``).
* ``tc_block_count``      -- number of 16x8 TC blocks a CSR matrix compacts to (the quantity being minimised), without
                             building the tiles.

TCA_reorder.py itself cannot run here (datasketch / cugraph / cudf / libMHCUDA are not in this image) and its result
depends on their hash seeds, so there is no bit-level parity to claim; ``oracle/tca_reorder.py`` restates its published
algorithm (exact candidates instead of LSH-approximate ones) for small graphs and ``tests/test_reorder.py`` compares TC-block
counts: ``cluster_reorder`` stays within 20 % of the restated TCA on small planted-partition graphs (3-4 % at 4 096 rows) (and beats it where
communities are larger than TCA's 0.2 similarity threshold can see).  Also tested: the permutation is valid, the product
is unchanged, and the TC-block count drops on a graph with planted communities whose labels were shuffled.
"""
from typing import Tuple

import torch

_P = (1 << 31) - 1   # Mersenne prime: (a * col + b) mod p stays inside int64 for 31-bit a, col


def _row_ids(indptr: torch.Tensor) -> torch.Tensor:
    M = indptr.numel() - 1
    deg = (indptr[1:] - indptr[:-1]).long()
    return torch.repeat_interleave(torch.arange(M, device=indptr.device), deg)


def minhash_signatures(indptr: torch.Tensor, indices: torch.Tensor, num_hashes: int = 2, seed: int = 0) -> torch.Tensor:
    """int64 [num_hashes, M]; rows without non-zeros get the sentinel p (they sort last)."""
    M = indptr.numel() - 1
    rows = _row_ids(indptr)
    cols = indices.long()
    g = torch.Generator().manual_seed(seed)
    coef = torch.randint(1, _P, (num_hashes, 2), generator=g, dtype=torch.int64)
    sig = torch.full((num_hashes, M), _P, dtype=torch.int64, device=indptr.device)
    for k in range(num_hashes):
        a, b = int(coef[k, 0]), int(coef[k, 1])
        h = (cols * a + b) % _P
        sig[k].scatter_reduce_(0, rows, h, reduce="amin", include_self=True)
    return sig


def lsh_reorder(indptr: torch.Tensor, indices: torch.Tensor, num_hashes: int = 2, seed: int = 0,
                levels: int = 4) -> torch.Tensor:
    """perm[i] = old id of the node placed at position i.

    Level 1 groups the rows by their min-hash NEIGHBOUR rep(r) (the neighbour with the smallest hash): all rows of a group
    share that neighbour.  A group is small, though, and consecutive groups are unrelated; so the groups are grouped in
    turn by the representative of their representative (rep(rep(r)) lives in the same neighbourhood), ``levels`` deep --
    a bottom-up clustering whose every step is one segmented min and one gather.  Sort key, most significant first:
    hash(rep^levels(r)), ..., hash(rep(r)), then the remaining min-hash values, then the node id (stable)."""
    M = indptr.numel() - 1
    dev = indptr.device
    sig = minhash_signatures(indptr, indices, num_hashes, seed)
    # node whose hash equals a row's first min-hash value: invert the (injective) hash through a sorted table
    g = torch.Generator().manual_seed(seed)
    coef = torch.randint(1, _P, (num_hashes, 2), generator=g, dtype=torch.int64)
    a, b = int(coef[0, 0]), int(coef[0, 1])
    node_hash = (torch.arange(M, device=dev, dtype=torch.int64) * a + b) % _P
    sorted_hash, by_hash = torch.sort(node_hash)
    has_nbr = sig[0] < _P
    pos = torch.searchsorted(sorted_hash, sig[0].clamp(max=_P - 1)).clamp(max=M - 1)
    rep = torch.where(has_nbr & (sorted_hash[pos] == sig[0]), by_hash[pos], torch.arange(M, device=dev))
    keys = [sig[k] for k in range(num_hashes - 1, 0, -1)]      # least significant first
    chain = rep
    keys.append(torch.where(has_nbr, node_hash[chain], torch.full_like(node_hash, _P)))
    for _ in range(levels - 1):
        chain = rep[chain]
        keys.append(torch.where(has_nbr, node_hash[chain], torch.full_like(node_hash, _P)))
    order = torch.arange(M, device=dev)
    for key in keys:                                            # LSD: stable sorts from the least significant key up
        order = order[torch.sort(key[order], stable=True).indices]
    return order


def _agglomerate(csig: torch.Tensor, alive: torch.Tensor, cap: int, need: int, max_rounds: int):
    """Rounds of proposer/acceptor matching over items with min-hash signatures ``csig`` [K, n] (see cluster_reorder).
    Returns (group id per item -- the id of a member item --, merged signatures valid at the group ids)."""
    K, n = csig.shape
    dev = csig.device
    group = torch.arange(n, device=dev)
    size = torch.ones(n, dtype=torch.int64, device=dev)
    csig = csig.clone()
    alive = alive.clone()
    roots = torch.nonzero(alive).flatten()
    idle = 0                                                               # consecutive rounds without a merge
    for rnd in range(max_rounds):
        if roots.numel() < 2 or idle >= 4:
            break
        idle += 1
        # 1. candidates: next root in each hash's sorted order, if the hash value is equal
        cand_a, cand_b = [], []
        for k in range(K):
            v = csig[k, roots]
            order = torch.argsort(v, stable=True)
            same = v[order][1:] == v[order][:-1]
            cand_a.append(roots[order[:-1]][same]); cand_b.append(roots[order[1:]][same])
        a = torch.cat(cand_a); b = torch.cat(cand_b)
        open_ = (size[a] < cap) & (size[b] < cap)           # TCA's rule: a cluster closes once it REACHES the cap
        a, b = a[open_], b[open_]
        if a.numel() == 0:
            continue
        pair = torch.unique(torch.minimum(a, b) * n + torch.maximum(a, b))
        a, b = pair // n, pair % n
        # 2. estimated Jaccard: agreeing signature components
        score = (csig[:, a] == csig[:, b]).sum(0)
        ok = score >= need
        a, b, score = a[ok], b[ok], score[ok]
        # 3. matching without chains: this round's items are split (by a hash of id and round) into proposers and
        #    acceptors; a proposer proposes to its best-scoring acceptor (ties -> smaller id), an acceptor takes its
        #    best-scoring proposer.  A pair only ever joins one proposer with one acceptor, so all accepted pairs merge.
        side = (((torch.arange(n, device=dev) * 2654435761 + (rnd + 1) * 40503) >> 9) & 1).bool()
        flip = side[a] & ~side[b]
        a, b = torch.where(flip, b, a), torch.where(flip, a, b)           # a = proposer side (False), b = acceptor side
        ok = ~side[a] & side[b]
        a, b, score = a[ok], b[ok], score[ok]
        if a.numel() == 0:
            continue
        best = torch.full((n,), -1, dtype=torch.int64, device=dev)
        best.scatter_reduce_(0, a, score * n + (n - 1 - b), reduce="amax", include_self=True)
        prop = torch.nonzero(best >= 0).flatten()                          # proposers with a candidate
        target = n - 1 - (best[prop] % n)
        acc = torch.full((n,), -1, dtype=torch.int64, device=dev)
        acc.scatter_reduce_(0, target, (best[prop] // n) * n + (n - 1 - prop), reduce="amax", include_self=True)
        hi = torch.nonzero(acc >= 0).flatten()                             # acceptors that received a proposal
        if hi.numel() == 0:
            continue
        lo = n - 1 - (acc[hi] % n)                                         # ... and the proposer each one takes
        idle = 0
        remap = torch.arange(n, device=dev)
        remap[hi] = lo
        group = remap[group]
        size[lo] += size[hi]
        csig[:, lo] = torch.minimum(csig[:, lo], csig[:, hi])
        alive[hi] = False
        roots = torch.nonzero(alive).flatten()
    return group, csig


def cluster_reorder(indptr: torch.Tensor, indices: torch.Tensor, window: int = 16, num_hashes: int = 64,
                    min_similarity: float = 0.08, max_rounds: int = 40, seed: int = 0, super_cap: int = 8) -> torch.Tensor:
    """Window-aware agglomerative clustering, the parallel counterpart of TCA_reorder.py's two greedy queues
    (third-party/DTC-SpMM/reordering/TCA_reorder.py:170-212: merge the most similar pair of row clusters, close a cluster
    at ``thres`` = 16 rows; :214-301: the same over the clusters, closed at 128).  perm[i] = old id of the node placed at
    position i.

    Every cluster carries a min-hash signature of the UNION of its rows' neighbour sets (the element-wise minimum of its
    members' signatures).  One round, all on the device:
      1. candidates: for each of the ``num_hashes`` hash functions, clusters are sorted by that signature value; neighbours
         in the sorted order with EQUAL values share the column that realises the minimum;
      2. a candidate pair is scored by the fraction of signature components that agree -- an unbiased estimate of the
         Jaccard similarity of the two neighbour sets;
      3. the clusters are split at random into proposers and acceptors; every proposer proposes to its best-scoring
         acceptor that is still open (a cluster closes once it reaches ``window`` rows, as in TCA), every acceptor takes
         its best proposer, and the accepted pairs merge (no chains, so no conflicts to resolve).
    Rounds repeat until four in a row merge nothing (clusters double at best per round: 16 rows need >= 4).  The same
    procedure then groups the clusters (``super_cap`` clusters per group, TCA's cache-aware level), and the groups are laid
    out in min-hash order of their signatures: rows of a cluster contiguous, clusters of a group contiguous.
    """
    M = indptr.numel() - 1
    dev = indptr.device
    sig = minhash_signatures(indptr, indices, num_hashes, seed)           # [K, M]; empty rows hold the sentinel
    K = num_hashes
    need = max(1, int(round(min_similarity * K)))
    nonempty = (indptr[1:] - indptr[:-1]) > 0                              # rows without non-zeros never merge
    cluster, csig = _agglomerate(sig, nonempty, window, need, max_rounds)
    # level 2: the clusters themselves, grouped the same way
    roots1 = torch.unique(cluster)
    c_index = torch.full((M,), -1, dtype=torch.int64, device=dev)
    c_index[roots1] = torch.arange(roots1.numel(), device=dev)
    sig1 = csig[:, roots1]
    group1, gsig = _agglomerate(sig1, sig1[0] < _P, super_cap, need, max_rounds)
    roots2 = torch.unique(group1)
    order2 = torch.arange(roots2.numel(), device=dev)
    for k in range(min(K, 4) - 1, -1, -1):                                 # LSD over the first few signature components
        order2 = order2[torch.sort(gsig[k, roots2[order2]], stable=True).indices]
    gpos = torch.empty(roots1.numel(), dtype=torch.int64, device=dev)
    gpos[roots2[order2]] = torch.arange(roots2.numel(), device=dev)
    n1 = roots1.numel()
    row_cluster = c_index[cluster]                                          # level-1 cluster index of every row
    key = (gpos[group1[row_cluster]] * n1 + row_cluster) * M + torch.arange(M, device=dev)
    return torch.sort(key).indices


def degree_reorder(indptr: torch.Tensor) -> torch.Tensor:
    deg = (indptr[1:] - indptr[:-1]).long()
    return torch.sort(deg, descending=True, stable=True).indices


def permute_graph(indptr: torch.Tensor, indices: torch.Tensor, perm: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``A' = P A P^T`` for a square CSR pattern: node ``perm[i]`` becomes node ``i``.  Columns stay sorted per row."""
    M = indptr.numel() - 1
    assert perm.numel() == M
    dev = indptr.device
    new_id = torch.empty(M, dtype=torch.int64, device=dev)
    new_id[perm] = torch.arange(M, device=dev)
    bits = max(1, int(M - 1).bit_length())
    keys = (new_id[_row_ids(indptr)] << bits) | new_id[indices.long()]
    keys = torch.sort(keys).values
    rows = keys >> bits
    cols = (keys & ((1 << bits) - 1)).to(torch.int32)
    new_indptr = torch.searchsorted(rows, torch.arange(M + 1, device=dev, dtype=rows.dtype)).to(torch.int32)
    return new_indptr, cols


def permute_rows(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """Rows of a dense [M, N] operand in the new node order (row i <- old row perm[i])."""
    return x[perm]


def unpermute_rows(y: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """Inverse of ``permute_rows``: result rows back in the original node order."""
    out = torch.empty_like(y)
    out[perm] = y
    return out


def tc_block_count(indptr: torch.Tensor, indices: torch.Tensor, blk_h: int = 16, blk_w: int = 8) -> int:
    """TC blocks after column compaction: sum over windows of ceil(distinct columns / blk_w), an edgeless window counting
    one block (the reference's rule, bmat_kernels.cuh:250-252,298-299)."""
    M = indptr.numel() - 1
    W = (M + blk_h - 1) // blk_h
    if W == 0:
        return 0
    bits = max(1, int(max(M, int(indices.max()) + 1 if indices.numel() else 1) - 1).bit_length())
    keys = torch.unique(((_row_ids(indptr) // blk_h) << bits) | indices.long())
    distinct = torch.bincount(keys >> bits, minlength=W)
    blocks = torch.clamp((distinct + blk_w - 1) // blk_w, min=1)
    return int(blocks.sum())
