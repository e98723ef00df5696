#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== small_blocks sweep"
for wl in "rmat25 0.25" "products 1.0" "amazon0505 1.0" "web-BerkStan 1.0" "DD 1.0" "ppi 1.0" "Yeast 1.0"; do
  set -- $wl
  for sb in 0 2 4 8 16 32; do
    echo "-- $1 small_blocks=$sb"; timeout -s KILL 300 python scripts/time_models.py --workload $1 --scale $2 --only 0/42/14,1/32/8 --small_blocks $sb 2>&1 | grep -v Warn | grep "sparse_rows\|model" | cut -c1-160
  done
done > $O/r2j_small_blocks_sweep.log 2>&1; grep -v "^R-MAT\|shaped" $O/r2j_small_blocks_sweep.log
