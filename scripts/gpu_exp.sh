#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/tc_probe.log
timeout 900 python scripts/time_models.py --workload reddit --only ${VARIANTS:-0/36/12,0/32/16,0/32/8} > gpurun_out/tm_exp.log 2>&1; echo "rc=$?"; grep -E "^model|Error|error" gpurun_out/tm_exp.log | tail -8
VAR=${PROFVAR:-0/36/12} bash scripts/gpu_prof_tc.sh
