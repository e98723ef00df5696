#!/bin/bash
# quick iteration: TC probe, spmm tests, model timing on reddit
mkdir -p gpurun_out
echo "== tc_probe"; timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tc_probe.log
echo "== spmm tests"; timeout 900 python -m pytest tests/test_spmm_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/t_spmm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_spmm.log
echo "== time_models reddit"; timeout 600 python scripts/time_models.py --workload reddit > gpurun_out/tm_reddit.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/tm_reddit.log
