#!/bin/bash
# pass 26: rows-per-warp CSR launch -- R-MAT-25 per-kernel breakdown, YeastH at N=512, parity tests of the CSR paths
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_spmm_gpu.py tests/test_parity_shapes_gpu.py -x -q -m gpu > gpurun_out/r2aa_t.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2aa_t.log
timeout -s KILL 600 python scripts/c5_breakdown_probe.py rmat25 1.0 2>&1 | grep -v "Warn\|warn" | tail -6 | cut -c1-180
for cfg in "512 fp16" "256 fp16" "128 fp32" "512 fp32"; do
  timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH $cfg 2>&1 | tail -1
done
