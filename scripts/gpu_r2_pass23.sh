#!/bin/bash
# pass 23: final one-GPU evidence on the final tree -- GPU suite, smoke, bench (both arms), C3 suite
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2x_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2x_t_gpu.log
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench reference arm"; timeout -s KILL 900 python bench.py --impl reference > $O/r2x_bench_ref_n1.json 2> $O/r2x_bench_ref_n1.err; echo "rc=$?"; cut -c1-400 $O/r2x_bench_ref_n1.json
echo "== bench"; timeout -s KILL 1200 python bench.py > $O/r2x_bench_n1.json 2> $O/r2x_bench_n1.err; echo "rc=$?"; cut -c1-700 $O/r2x_bench_n1.json
echo "== C3 suite"; timeout -s KILL 1500 python scripts/suite.py --out $O/r2x_suite_c3.csv > $O/r2x_suite_c3.log 2>&1; echo "rc=$?"; tail -4 $O/r2x_suite_c3.log
