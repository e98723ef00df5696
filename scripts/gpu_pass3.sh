#!/bin/bash
mkdir -p gpurun_out
echo "== spmm tests"; timeout 600 python -m pytest tests/test_spmm_gpu.py tests/test_ref_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/t_spmm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_spmm.log
echo "== variants"; timeout 300 python scripts/time_models.py --workload reddit --only ${VARIANTS:-0/36/12,0/40/16,0/32/8} > gpurun_out/tm_reddit3.log 2>&1; echo "rc=$?"; grep -E "^model|M=|rror" gpurun_out/tm_reddit3.log
