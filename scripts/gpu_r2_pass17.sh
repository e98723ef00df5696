#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2r_t_gpu.log 2>&1; echo "rc=$?"; tail -5 $O/r2r_t_gpu.log | grep -v "Warning\|sparse_csr\|^$"
echo "== C3 suite"; timeout -s KILL 1500 python scripts/suite.py --out $O/r2r_suite_c3.csv > $O/r2r_suite_c3.log 2>&1; echo "rc=$?"; tail -2 $O/r2r_suite_c3.log
echo "== bench"; timeout -s KILL 1200 python bench.py > $O/r2r_bench_n1.json 2> $O/r2r_bench_n1.err; echo "rc=$?"; grep "timed loop" $O/r2r_bench_n1.err | cut -c1-120
