"""TEST INFRASTRUCTURE ONLY -- a CPU restatement of the row reordering the reference's published numbers depend on.

The reference benchmarks Voltrix on ``*.reorder.npz`` graphs (bench/bench_all.py:21-23,120-149), relabelled offline by
third-party/DTC-SpMM/reordering/TCA_reorder.py.  That script needs datasketch, cugraph, cudf and libMHCUDA, none of which is
in this image, so it cannot be run here: PARITY UNPINNED -- this file restates its published algorithm and is used to compare
TC-block counts with ``voltrix.reorder`` on small graphs, not to claim identical permutations.

What TCA_reorder.py does (line numbers of that file):
  1. candidate pairs: MinHash-LSH query per row, threshold 0.2 (:24,55-87,140-166), scored by exact Jaccard similarity of
     the neighbour sets (cugraph.jaccard, :155).  Restated with EXACT candidates: every pair of rows with Jaccard >= 0.2
     (what the LSH approximates) -- computable for a few thousand rows.
  2. "TCU-aware" clustering (:170-212): pairs in a max-priority queue by similarity; union two roots (smaller into larger);
     a cluster that reaches ``thres`` = 16 rows (the window height) is closed; a popped pair whose ends are not both roots
     is re-queued as (root, root) with the Jaccard of the ROOT ROWS' neighbour lists (:196-203 -- of the representative
     rows, not of the merged sets).
  3. "cache-aware" clustering of the clusters (:214-301): the same procedure on the union neighbour sets of the clusters,
     threshold 0.2, closed at ``c_thres`` = 128 clusters.
  4. new order = clusters of clusters -> clusters -> rows, in dict insertion order (:306-311).
"""
import heapq
from typing import List

import numpy as np


def _jaccard(a: set, b: set) -> float:
    if not a or not b:
        return 0.0
    return len(a & b) / len(a | b)


def _candidates(sets: List[set], thres: float):
    """All pairs (i < j) with Jaccard >= thres, via an inverted index (exact stand-in for the LSH query)."""
    inv = {}
    for i, s in enumerate(sets):
        for c in s:
            inv.setdefault(c, []).append(i)
    seen = set()
    out = []
    for members in inv.values():
        if len(members) > 2000:      # a hub column pairs everything with everything; such pairs score far below thres
            continue
        for x in range(len(members)):
            for y in range(x + 1, len(members)):
                p = (members[x], members[y])
                if p in seen:
                    continue
                seen.add(p)
                sim = _jaccard(sets[p[0]], sets[p[1]])
                if sim >= thres:
                    out.append((sim, p[0], p[1]))
    return out


def _greedy_cluster(sets: List[set], thres: float, cap: int) -> List[List[int]]:
    n = len(sets)
    parent = list(range(n))
    size = [1] * n
    closed = [False] * n
    live = n

    def root(i):
        while i != parent[i]:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    heap = []
    queued = set()
    for sim, a, b in _candidates(sets, thres):
        heapq.heappush(heap, (-sim, a, b))
        queued.add((a, b))
    while heap and live > 0:
        _, p1, p2 = heapq.heappop(heap)
        queued.discard((min(p1, p2), max(p1, p2)))
        if p1 == parent[p1] and p2 == parent[p2]:
            if closed[p1] or closed[p2]:
                continue
            small, big = (p1, p2) if size[p1] < size[p2] else (p2, p1)
            parent[small] = big
            live -= 1
            size[big] += size[small]
            if size[big] >= cap:
                closed[big] = True
                live -= 1
        else:
            r1, r2 = root(p1), root(p2)
            if closed[r1] or closed[r2]:
                continue
            key = (min(r1, r2), max(r1, r2))
            if r1 != r2 and key not in queued:
                heapq.heappush(heap, (-_jaccard(sets[r1], sets[r2]), r1, r2))
                queued.add(key)
    clusters = {}
    for i in range(n):
        clusters.setdefault(root(i), []).append(i)
    return list(clusters.values())


def tca_reorder(indptr: np.ndarray, indices: np.ndarray, thres: int = 16, lsh_thres: float = 0.2, c_thres: int = 128,
                cluster_thres: float = 0.2) -> np.ndarray:
    """Permutation (new position -> old row id) of TCA_reorder.py's two-level clustering, for small matrices."""
    M = indptr.size - 1
    rows = [set(indices[indptr[i]:indptr[i + 1]].tolist()) for i in range(M)]
    level1 = _greedy_cluster(rows, lsh_thres, thres)
    unions = [set().union(*[rows[r] for r in c]) for c in level1]
    level2 = _greedy_cluster(unions, cluster_thres, c_thres)
    order = [r for group in level2 for k in group for r in level1[k]]
    assert sorted(order) == list(range(M))
    return np.asarray(order, dtype=np.int64)
