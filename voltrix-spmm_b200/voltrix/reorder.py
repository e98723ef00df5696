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
* ``degree_reorder``      -- permutation by descending degree (cheap baseline; groups the hubs).
* ``permute_graph``       -- relabel a square adjacency matrix, rows AND columns: ``A' = P A P^T``.  Then
                             ``A' (P B) = P (A B)``: permute the rows of B with ``perm`` going in and read row ``i`` of the
                             result as node ``perm[i]`` (helpers ``permute_rows`` / ``unpermute_rows``).
* ``tc_block_count``      -- number of 16x8 TC blocks a CSR matrix compacts to (the quantity being minimised), without
                             building the tiles.

No parity oracle exists for the reference's reordering (datasketch / cugraph are not in this image and the result
depends on their hash seeds); what is tested is that the permutation is valid, that the product is unchanged, and that
the TC-block count drops on a graph with planted communities whose labels were shuffled.
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
