"""Multi-GPU host logic on CPU: partitioning, shard extraction, broadcast / all-gather plumbing under gloo
(world_size 2).  The local SpMM is injected (the oracle's CPU SpMM) -- the CUDA path itself is covered by the
-m gpu tests; here the point is that shards + collectives reproduce the single-process result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_partition_rows_balanced_and_aligned():
    from voltrix.distributed import partition_rows
    rng = np.random.default_rng(0)
    deg = torch.from_numpy((rng.pareto(1.5, 10_007) * 20).astype(np.int64))
    for world in (1, 2, 3, 4, 8):
        ranges = partition_rows(deg, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == deg.numel()
        for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
            assert b == c and a % 16 == 0 and b % 16 == 0 and a <= b
        loads = [int(deg[a:b].sum()) for a, b in ranges]
        # no shard is heavier than the ideal share plus one window's worth of weight
        max_window = int(torch.nn.functional.pad(deg, (0, 16 - deg.numel() % 16)).view(-1, 16).sum(1).max())
        assert max(loads) <= sum(loads) / world + max_window


def test_partition_rows_skewed_hub():
    from voltrix.distributed import partition_rows
    deg = torch.ones(1600, dtype=torch.int64)
    deg[5] = 100_000            # one hub row outweighs everything else
    ranges = partition_rows(deg, 4)
    assert ranges[0] == (0, 16)                      # the hub's window is a shard of its own
    assert all(a <= b for a, b in ranges) and ranges[-1][1] == 1600
    loads = [int(deg[a:b].sum()) for a, b in ranges]
    assert max(loads) == loads[0] == 100_015          # bottleneck = the hub window, nothing else piles onto it


def test_partition_more_ranks_than_windows():
    from voltrix.distributed import partition_rows
    ranges = partition_rows(torch.ones(20, dtype=torch.int64), 8)
    assert ranges[0][0] == 0 and ranges[-1][1] == 20 and all(a <= b for a, b in ranges)
    assert sum(b - a for a, b in ranges) == 20


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from voltrix.distributed import ShardedSpMM
        rng = np.random.default_rng(5)
        import scipy.sparse as sp
        M, N = 1000, 24     # M % 16 != 0
        A = sp.random(M, M, density=0.03, format="csr", random_state=rng)
        indptr = torch.from_numpy(A.indptr.astype(np.int32))
        indices = torch.from_numpy(A.indices.astype(np.int32))

        def local_preprocess(ip, ix, rows, cols):
            return (ip.numpy(), ix.numpy())

        def local_spmm(state, feat):
            ip, ix = state
            return torch.from_numpy(oracle.c().spmm_csr(ip, ix, feat.numpy()))

        sh = ShardedSpMM(indptr, indices, M, local_preprocess=local_preprocess, local_spmm=local_spmm)
        # B lives on rank 0 only; the broadcast is the path's one exchange step
        B = torch.from_numpy(np.random.default_rng(9).standard_normal((M, N)).astype(np.float32)) if rank == 0 \
            else torch.zeros(M, N)
        sh.broadcast_features(B, src=0)
        c_local = sh.spmm(B)
        assert c_local.shape == (sh.r1 - sh.r0, N)
        full = sh.all_gather(c_local)
        want = torch.from_numpy(oracle.c().spmm_csr(A.indptr.astype(np.int32), A.indices.astype(np.int32), B.numpy()))
        ok = torch.equal(full, want) and sh.ranges[0][0] == 0 and sh.ranges[-1][1] == M
        # reusing a caller-owned result buffer gives the same rows
        again = sh.all_gather(c_local, out=torch.full((M, N), float("nan")))
        ok = ok and torch.equal(again, want)
        # per-step operand exchange: every rank owns only ITS row slice of the host operand (the rest is poisoned);
        # upload + all-gather must rebuild the whole operand on every rank, padding rows excluded
        from voltrix.distributed import operand_slices, upload_slice_and_all_gather
        chunk, slices = operand_slices(M, world)
        lo, hi = slices[rank]
        mine_only = torch.full((M, N), float("nan"))
        mine_only[lo:hi] = B[lo:hi]
        buf = torch.full((chunk * world, N), float("nan"))
        got = upload_slice_and_all_gather(buf, mine_only, rank, world)
        ok = ok and got.shape == (M, N) and torch.equal(got, B)
        # the local_csr route (no rank holds the whole matrix) cuts the same shards
        def local_csr(r0, r1):
            lo_e, hi_e = int(indptr[r0]), int(indptr[r1])
            return (indptr[r0:r1 + 1] - lo_e).to(torch.int32), indices[lo_e:hi_e]
        sh2 = ShardedSpMM(None, None, M, weights=(indptr[1:] - indptr[:-1]) + 12, local_preprocess=local_preprocess,
                          local_spmm=local_spmm, local_csr=local_csr)
        ok = ok and sh2.ranges == sh.ranges and torch.equal(sh2.spmm(B), c_local)
        q.put((rank, bool(ok), sh.ranges))
    finally:
        dist.destroy_process_group()


def test_sharded_spmm_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    assert results[0][2] == results[1][2], "every rank must compute the same partition"


def test_row_cost_balances_skewed_row_counts():
    """R-MAT-like skew: a dense head (few rows, many non-zeros each) and a long sparse tail.  Equal-nnz shards give the
    tail rank 6x the rows of the head rank; with the per-row cost (ROW_COST non-zeros per row: its share of the window's
    work item and the C row write, measured on the 2-GPU R-MAT run) the modelled cost nnz + ROW_COST * rows is balanced."""
    import torch
    from voltrix.distributed import ROW_COST, partition_rows
    g = torch.Generator().manual_seed(0)
    head = torch.randint(80, 120, (30_000,), generator=g)
    tail = torch.randint(0, 4, (1_500_000,), generator=g)
    deg = torch.cat([head, tail])

    def costs(ranges):
        return [int(deg[a:b].sum()) + ROW_COST * (b - a) for a, b in ranges]

    by_nnz = costs(partition_rows(deg, 4))
    by_cost = costs(partition_rows(deg + ROW_COST, 4))
    assert max(by_cost) / (sum(by_cost) / 4) < 1.01
    assert max(by_nnz) / (sum(by_nnz) / 4) > 1.5
    rng = partition_rows(deg + ROW_COST, 4)
    assert rng[0][0] == 0 and rng[-1][1] == deg.numel() and all(a % 16 == 0 for a, _ in rng)
    assert all(rng[k][1] == rng[k + 1][0] for k in range(3))


def test_operand_slices_cover_rows_once():
    from voltrix.distributed import operand_slices
    for rows, world in ((1000, 2), (1000, 8), (7, 8), (232_965, 8), (16, 1), (0, 4)):
        chunk, slices = operand_slices(rows, world)
        assert len(slices) == world and chunk * world >= rows
        assert slices[0][0] == 0 and slices[-1][1] == rows
        assert all(a <= b and b - a <= chunk for a, b in slices)
        assert all(slices[k][1] == slices[k + 1][0] for k in range(world - 1))
        assert all(a == min(k * chunk, rows) for k, (a, _) in enumerate(slices))   # slice k starts at rank k's all-gather slot
