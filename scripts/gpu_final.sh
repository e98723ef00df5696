#!/bin/bash
# Round-end evidence pass on one GPU.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_final.json
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_final_ref.json
echo "== reorder"; timeout 300 python scripts/reorder_probe.py > gpurun_out/reorder.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/reorder.log
echo "== suite"; timeout 900 python scripts/suite.py --out gpurun_out/suite_final.csv > gpurun_out/suite_final.log 2>&1; echo "rc=$?"; grep -c . gpurun_out/suite_final.csv
for v in 36/12 42/14; do
  echo "== ncu full $v"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o gpurun_out/prof_tc_final_${v/\//_} \
    python scripts/time_models.py --workload reddit --only 0/$v --once > gpurun_out/ncu_tc_$$.log 2>&1; echo "rc=$?"
done
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-baselines > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
