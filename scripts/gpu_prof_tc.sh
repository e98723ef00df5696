#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o gpurun_out/prof_tc \
    python scripts/time_models.py --workload ${WL:-reddit} --only ${VAR:-0/32/8} --once > gpurun_out/ncu_tc.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_tc.log
