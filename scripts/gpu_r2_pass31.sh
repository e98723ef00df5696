#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2af_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2af_t_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py --extra "" > $O/r2af_bench_n1.json 2> $O/r2af_bench_n1.err; echo "rc=$?"; cut -c1-200 $O/r2af_bench_n1.json
