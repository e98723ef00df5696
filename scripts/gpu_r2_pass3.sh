#!/bin/bash
# Round 2, third one-GPU pass: GPU tests (model 4, autograd, fast path), fp32 variants on the Reddit-shaped graph, C3 suite.
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2c_t_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/r2c_t_gpu.log | grep -v Warning
echo "== fp32 variants reddit"; timeout -s KILL 400 python scripts/time_models.py --workload reddit --dtype fp32 --only 4/24/8,3/24/8,1/32/8 > $O/r2c_tm_reddit_fp32.log 2>&1; echo "rc=$?"; tail -6 $O/r2c_tm_reddit_fp32.log
echo "== C3 suite"; timeout -s KILL 1200 python scripts/suite.py --out $O/r2c_suite_c3.csv > $O/r2c_suite_c3.log 2>&1; echo "rc=$?"; tail -2 $O/r2c_suite_c3.log
echo "== bench (no extras)"; timeout -s KILL 600 python bench.py --extra "" > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "rc=$?"; cut -c1-250 $O/r2c_bench_n1.json
