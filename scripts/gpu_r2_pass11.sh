#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2k_t_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/r2k_t_gpu.log | grep -v "Warning\|sparse_csr\|^$"
echo "== small_blocks with the group-per-row kernel on the sparse rows"
for wl in "rmat25 0.25" "Yeast 1.0" "DD 1.0" "amazon0505 1.0"; do
  set -- $wl
  for sb in 0 8 16 32 64; do
    echo "-- $1 small_blocks=$sb"; timeout -s KILL 300 python scripts/time_models.py --workload $1 --scale $2 --only 0/42/14,1/32/8 --small_blocks $sb 2>&1 | grep -v Warn | grep "sparse_rows\|model" | cut -c1-200
  done
done > $O/r2k_small_blocks_sweep.log 2>&1; grep -v "max scaled" $O/r2k_small_blocks_sweep.log | cut -c1-175
echo "== bench"; timeout -s KILL 1200 python bench.py > $O/r2k_bench_n1.json 2> $O/r2k_bench_n1.err; echo "rc=$?"; grep "timed loop\|ok'" $O/r2k_bench_n1.err | cut -c1-200
