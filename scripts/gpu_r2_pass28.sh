#!/bin/bash
# pass 28: rows-per-warp on the warp-per-row kernel only -- GPU suite, then RPW 2 / 4 / 8 on R-MAT-25's sparse rows and YeastH
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2ac_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2ac_t_gpu.log
export VOLTRIX_EXTRA_NVCC_FLAGS=-DVX_CSR_RPW_PROBE
for rpw in 2 4 8; do
  echo "== RPW $rpw"
  VX_CSR_RPW=$rpw timeout -s KILL 600 python scripts/c5_breakdown_probe.py rmat25 1.0 2>&1 | grep "csr_rows" | cut -c1-160
  VX_CSR_RPW=$rpw timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH 512 fp16 2>&1 | tail -1 | cut -c1-120
  VX_CSR_RPW=$rpw timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH 128 fp32 2>&1 | tail -1 | cut -c1-120
done
