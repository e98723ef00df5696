#!/bin/bash
# pass 27: rows-per-warp on both CSR kernels -- full GPU suite (new at-scale short-row test), YeastH probes, sparse end of C3
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2ab_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2ab_t_gpu.log
for cfg in "32 fp16" "128 fp16" "512 fp16" "128 fp32"; do
  timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH $cfg 2>&1 | tail -1
done
timeout -s KILL 1500 python scripts/suite.py --out $O/r2ab_suite_c3.csv > $O/r2ab_suite_c3.log 2>&1; echo "rc=$?"
cut -d, -f1,5,8,9,10,11,14 $O/r2ab_suite_c3.csv | grep -i "yeast\|DD\|com-amazon\|protein\|dataset"
