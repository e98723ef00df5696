"""Per-kernel time of one SpMM step on a large workload (default: the full R-MAT-25 configuration), from the torch profiler:
how the step splits between the tcgen05 kernel, the CUDA-core rows of the sparse windows and the fix-up pass."""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
import voltrix  # noqa: E402
import bench as B  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "rmat25"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda")
indptr, indices, N, desc = B.make_workload(name, dev, scale)
M, nnz = indptr.numel() - 1, indices.numel()
st = voltrix.csr_preprocess(indptr, indices, M)
plan = st[1]._vx_plan
deg = (indptr[1:] - indptr[:-1])
srows = plan.sparse_rows[: plan.num_sparse_rows].long()
sdeg = deg[srows]
print(f"{desc}: M={M} nnz={nnz} TCB={plan.total_blocks} items={plan.num_items} fixups={plan.num_fixups} "
      f"sparse_rows={plan.num_sparse_rows} (nnz {int(sdeg.sum())}, mean degree {float(sdeg.float().mean()):.2f}, "
      f"empty {int((sdeg == 0).sum())}) cap={plan.cap}", flush=True)
del indices
feat = torch.rand(M, N, device=dev).half()
out = torch.empty(M, N, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
fn = lambda: voltrix.spmm(*st, M, nnz, feat, out=out)   # noqa: E731
fn(); torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        flush.zero_(); fn()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total / max(e.count, 1) > 5:
        print(f"  {e.key[:110]:110s} x{e.count}  {e.device_time_total / e.count / 1e3:9.3f} ms", flush=True)
