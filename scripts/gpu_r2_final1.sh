#!/bin/bash
# Round-2 final one-GPU evidence: ncu --set full on the shipped variants (C2 14/7, C4 15/5, R-MAT-23 22/11), launch list.
O=gpurun_out; mkdir -p $O
for spec in "reddit 1.0 0/14/7" "products 1.0 0/15/5" "rmat25 0.25 0/22/11"; do
  set -- $spec
  echo "== ncu full $1 $3"; timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o $O/r2p_prof_tc_$1 \
    python scripts/time_models.py --workload $1 --scale $2 --only $3 --once > $O/r2p_ncu_$1.log 2>&1; echo "rc=$?"
done
echo "== ncu launch list"; timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2p_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-baselines --extra "" > $O/r2p_bench_ncu.log 2>&1; echo "rc=$?"
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
