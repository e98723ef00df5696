"""cuSPARSE baseline (reference: bench/bm_sparse.py:1-52): torch.sparse_csr @ dense, 10 warm + 100 timed
iterations, CUDA events; prints `[cuSPARSE] Elapsed time: x ms`."""
import numpy as np
import torch

indices = torch.tensor(np.loadtxt("indices.csv", delimiter=",", dtype=np.int32), dtype=torch.int32).cuda()
offsets = torch.tensor(np.loadtxt("indptr.csv", delimiter=",", dtype=np.int32), dtype=torch.int32).cuda()
N = offsets.numel() - 1
csr = torch.sparse_csr_tensor(offsets, indices, values=torch.ones_like(indices).float(), size=(N, N)).cuda()
print(indices.numel())
weight = torch.tensor(np.fromfile("feat.csv", dtype=np.float32)).cuda().view(N, -1)


def f():
    return csr @ weight


iters = 100
for _ in range(10):
    o = f()
torch.cuda.synchronize()
start_event = torch.cuda.Event(enable_timing=True)
end_event = torch.cuda.Event(enable_timing=True)
start_event.record()
for _ in range(iters):
    o = f()
end_event.record()
torch.cuda.synchronize()
o = f()
o_base = np.fromfile("output_base.csv", dtype=np.float32).reshape(*list(o.shape))
print(np.allclose(o.detach().cpu().numpy(), o_base, atol=1e-1))
print(f"[cuSPARSE] Elapsed time: {start_event.elapsed_time(end_event) / iters:.4f} ms")
