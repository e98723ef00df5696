#!/bin/bash
# Round 2, fourth one-GPU pass: GPU tests (model 4 with carrier scaling), fp32 variants, C3 suite, competitors on two graphs.
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2d_t_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/r2d_t_gpu.log | grep -v "Warning\|sparse_csr\|^$"
echo "== fp32 variants reddit"; timeout -s KILL 400 python scripts/time_models.py --workload reddit --dtype fp32 --only 4/24/8,3/24/8,1/32/8 > $O/r2d_tm_reddit_fp32.log 2>&1; echo "rc=$?"; tail -6 $O/r2d_tm_reddit_fp32.log
echo "== C3 suite"; timeout -s KILL 1200 python scripts/suite.py --out $O/r2d_suite_c3.csv > $O/r2d_suite_c3.log 2>&1; echo "rc=$?"; tail -2 $O/r2d_suite_c3.log
echo "== competitors (bench_all on ddi, amazon0505; N = 128, 256, 512)"
cd bench && timeout -s KILL 1500 python bench_all.py --datasets ddi amazon0505 FraudYelp-RSR --feature_dims 128 256 512 --results ../$O/r2d_competitors.csv > ../$O/r2d_competitors.log 2>&1; echo "rc=$?"; cd ..; tail -40 $O/r2d_competitors.log
