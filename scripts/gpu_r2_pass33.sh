#!/bin/bash
# pass 33: size-dependent small-window rule -- GPU suite, C3 suite, bench with the C4 / C5 extras
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2ah_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2ah_t_gpu.log
timeout -s KILL 900 python scripts/suite.py --out $O/r2ah_suite_c3.csv > $O/r2ah_suite_c3.log 2>&1; echo "rc=$?"
timeout -s KILL 900 python bench.py > $O/r2ah_bench_n1.json 2> $O/r2ah_bench_n1.err; echo "rc=$?"; cut -c1-160 $O/r2ah_bench_n1.json
grep "timed loop done" $O/r2ah_bench_n1.err
