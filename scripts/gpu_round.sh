#!/bin/bash
# First-contact GPU script: probe, tests, each under its own timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== tc_probe"; timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/tc_probe.log
echo "== preprocess tests"; timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -m gpu --timeout 300 > gpurun_out/t_pre.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_pre.log
echo "== spmm tests"; timeout 900 python -m pytest tests/test_spmm_gpu.py -q -m gpu --timeout 300 > gpurun_out/t_spmm.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_spmm.log
echo "== other gpu tests"; timeout 600 python -m pytest tests/test_ref_gpu.py tests/test_jit.py tests/test_distributed.py -q -m gpu --timeout 300 > gpurun_out/t_other.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_other.log
