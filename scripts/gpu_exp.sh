#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/time_models.py --workload reddit --only 0/36/12,0/24/12,0/32/16,0/48/16,0/32/8 > gpurun_out/tm_exp.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/tm_exp.log
VAR=0/36/12 bash scripts/gpu_prof_tc.sh
