"""Single-process replay of one rank's shard of bench.py (no NCCL): preprocess rows of `--rank` of `--world`, then the timed
loop and the serial e2e loop with a progress line per step -- to tell a slow step from a hung one."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
import bench as B  # noqa: E402
from voltrix.distributed import partition_rows, shard_csr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="rmat25")
ap.add_argument("--scale", type=float, default=0.0625)
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--rank", type=int, default=1)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--model", type=int, default=None)
args = ap.parse_args()
dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload(args.workload, dev, args.scale)
M = indptr.numel() - 1
r0, r1 = partition_rows(indptr[1:] - indptr[:-1], args.world)[args.rank]
lp, li = shard_csr(indptr, indices, r0, r1)
rows, nnz = r1 - r0, li.numel()
blk, packed, hind = voltrix.csr_preprocess(lp, li, rows, num_cols=M)
plan = packed._vx_plan
print(f"rows [{r0},{r1}) nnz={nnz} TCB={plan.total_blocks} items={plan.num_items} sparse_rows={plan.num_sparse_rows} "
      f"fixups={plan.num_fixups} slots={plan.num_slots} cap={plan.cap}", flush=True)
feat = torch.rand(M, N, device=dev).half()
out = torch.empty(rows, N, device=dev)


def step(f):
    if args.model is None:
        return voltrix.spmm(blk, packed, hind, rows, nnz, f, out=out)
    voltrix.spmm_kernel(blk, packed, hind, num_nodes=rows, num_edges=nnz, embedding_dim=N, input=f, output=out,
                        model=args.model, stages=36, npw=12)
    return out


t0 = time.perf_counter(); step(feat); torch.cuda.synchronize()
print(f"first call (tune) {time.perf_counter() - t0:.2f}s tuned={list(voltrix.jit_tuner.tuned_keys.values())}", flush=True)
for i in range(args.steps):
    t0 = time.perf_counter(); step(feat); torch.cuda.synchronize()
    print(f"device step {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
feat_host = feat.cpu().pin_memory()
out_host = torch.empty(rows, N).pin_memory()
feat_dev = torch.empty_like(feat)
for i in range(args.steps):
    t0 = time.perf_counter()
    feat_dev.copy_(feat_host, non_blocking=True)
    o = step(feat_dev)
    out_host.copy_(o, non_blocking=True)
    torch.cuda.synchronize()
    print(f"serial e2e step {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
print("done", flush=True)
