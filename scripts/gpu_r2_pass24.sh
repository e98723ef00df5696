#!/bin/bash
# final one-GPU evidence: GPU suite, smoke, C3 suite, fp32 repeatability, both bench arms
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2zz_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2zz_t_gpu.log
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== C3 suite"; timeout -s KILL 1500 python scripts/suite.py --out $O/r2zz_suite_c3.csv > $O/r2zz_suite_c3.log 2>&1; echo "rc=$?"; grep "^reddit" $O/r2zz_suite_c3.log | tail -5
echo "== reddit fp32 N=512 / N=256, repeated (fresh operand each time)"
timeout -s KILL 600 python - <<'PY' 2>&1 | tail -12
import sys, os, torch, numpy as np
sys.path.insert(0, "voltrix-spmm_b200"); sys.path.insert(0, ".")
import voltrix, bench as B
dev = torch.device("cuda")
indptr, indices, _, _ = B.make_workload("reddit", dev, 1.0)
M, nnz = indptr.numel() - 1, indices.numel()
st = voltrix.csr_preprocess(indptr, indices, M)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
for N in (128, 256, 512):
    out = torch.empty(M, N, device=dev)
    res = []
    for rep in range(6):
        feat = torch.randn(M, N, device=dev) if rep % 2 else torch.rand(M, N, device=dev)
        voltrix.spmm(*st, M, nnz, feat, out=out); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); voltrix.spmm(*st, M, nnz, feat, out=out); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        res.append(round(float(np.median(ts)), 3))
    print(f"reddit fp32 N={N}: ms per operand (uniform, normal alternating) {res}", flush=True)
PY
echo "== bench reference arm"; timeout -s KILL 900 python bench.py --impl reference > $O/r2zz_bench_ref_n1.json 2> $O/r2zz_bench_ref_n1.err; echo "rc=$?"
echo "== bench"; timeout -s KILL 1200 python bench.py > $O/r2zz_bench_n1.json 2> $O/r2zz_bench_n1.err; echo "rc=$?"; cut -c1-300 $O/r2zz_bench_n1.json
