#!/bin/bash
# Round 2, second one-GPU pass: all GPU tests (weighted tensor-core path, chain cap), bench, weighted probe, full C3 suite,
# ncu on the full R-MAT.
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2b_t_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/r2b_t_gpu.log
echo "== bench"; timeout -s KILL 1200 python bench.py > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err; echo "rc=$?"; cut -c1-300 $O/r2b_bench_n1.json; grep "parity\|timed loop" $O/r2b_bench_n1.err
echo "== weighted probe"; for wl in reddit products; do timeout -s KILL 400 python scripts/weighted_probe.py --workload $wl > $O/r2b_weighted_$wl.log 2>&1; echo "rc=$?"; cat $O/r2b_weighted_$wl.log | tail -12; done
echo "== C3 suite"; timeout -s KILL 1200 python scripts/suite.py --out $O/r2b_suite_c3.csv > $O/r2b_suite_c3.log 2>&1; echo "rc=$?"; tail -3 $O/r2b_suite_c3.log
echo "== ncu full rmat25"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o $O/r2b_prof_tc_rmat25 \
    python scripts/time_models.py --workload rmat25 --only 0/42/14 --once > $O/r2b_ncu_rmat25.log 2>&1; echo "rc=$?"
