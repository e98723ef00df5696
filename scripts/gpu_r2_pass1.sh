#!/bin/bash
# Round 2, first GPU pass (one B200): smoke gate, every GPU test, bench with the extra workloads, N sweep on the
# Reddit-shaped graph, ncu --set full on the HBM-bound configurations, L2-policy probe on the R-MAT.
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; rc=$?; tail -2 $O/r2_smoke.log
if [ $rc -ne 0 ]; then
  echo "SMOKE FAILED rc=$rc -- running one small test for the trace and stopping"
  timeout -s KILL 300 python -m pytest tests/test_spmm_gpu.py -x -q -m gpu --timeout 120 -k "test_every_path_matches_oracle and m1000_sparse and 128" 2>&1 | tail -40
  exit 1
fi
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2_t_gpu.log 2>&1; echo "rc=$?"; tail -15 $O/r2_t_gpu.log
echo "== bench (headline + products + rmat25 extras)"; timeout -s KILL 1200 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "rc=$?"; cut -c1-600 $O/r2_bench_n1.json; tail -5 $O/r2_bench_n1.err
echo "== suite reddit N sweep"; timeout -s KILL 400 python scripts/suite.py --datasets reddit --feature_dims 128 256 512 --out $O/r2_suite_reddit.csv > $O/r2_suite_reddit.log 2>&1; echo "rc=$?"; cat $O/r2_suite_reddit.csv
echo "== variants on products / rmat23"
timeout -s KILL 300 python scripts/time_models.py --workload products --only 0/36/12,0/42/14,1/32/8 > $O/r2_tm_products.log 2>&1; tail -4 $O/r2_tm_products.log
timeout -s KILL 300 python scripts/time_models.py --workload rmat25 --scale 0.25 --only 0/36/12,0/42/14,1/32/8 > $O/r2_tm_rmat23.log 2>&1; tail -4 $O/r2_tm_rmat23.log
echo "== L2 policy probe (R-MAT 23): hub rows evict_last, others evict_first"
timeout -s KILL 300 python scripts/isolate.py --workload rmat25 --scale 0.25 --variant 0/42/14 --flags="|-DVX_TC_HUB_POPC=5|-DVX_TC_HUB_POPC=6" > $O/r2_hubpol_rmat23.log 2>&1; tail -4 $O/r2_hubpol_rmat23.log
for wl in "products 1.0" "rmat25 0.25" "reddit 1.0"; do
  set -- $wl
  echo "== ncu full $1"; timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o $O/r2_prof_tc_$1 \
    python scripts/time_models.py --workload $1 --scale $2 --only 0/42/14 --once > $O/r2_ncu_$1.log 2>&1; echo "rc=$?"
done
echo "== ncu launch list"; timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-baselines --extra "" > $O/r2_bench_ncu.log 2>&1; echo "rc=$?"
ls -la $O | grep r2_
