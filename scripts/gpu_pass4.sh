#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gpu.log
echo "== reddit fp32"; timeout 300 python scripts/time_models.py --workload reddit --dtype fp32 > gpurun_out/tm_reddit_fp32.log 2>&1; echo "rc=$?"; grep -E "^model|M=|rror|diff" gpurun_out/tm_reddit_fp32.log
echo "== suite subset"; timeout 600 python scripts/suite.py --datasets Yeast DD FraudYelp-RSR ppi --out gpurun_out/suite_r1c.csv > gpurun_out/suite2.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/suite2.log | tail -22 | cut -c1-160
