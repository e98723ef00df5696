#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== sparse_ratio sweep"
for wl in "rmat25 0.25" "products 1.0" "amazon0505 1.0" "web-BerkStan 1.0" "FraudYelp-RSR 1.0"; do
  set -- $wl
  for sr in 0.0 0.25 0.5 1.0 2.0; do
    echo "-- $1 sparse_ratio=$sr"; timeout -s KILL 300 python scripts/time_models.py --workload $1 --scale $2 --only 0/42/14,0/40/24 --sparse_ratio $sr 2>&1 | grep -v Warn | grep "sparse_rows\|model" | cut -c1-150
  done
done > $O/r2i_sparse_ratio_sweep.log 2>&1; cat $O/r2i_sparse_ratio_sweep.log
echo "== ncu reddit N=512"; timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o $O/r2i_prof_tc_reddit_n512 \
    python scripts/time_models.py --workload reddit --N 512 --only 0/42/14 --once > $O/r2i_ncu_reddit_n512.log 2>&1; echo "rc=$?"
