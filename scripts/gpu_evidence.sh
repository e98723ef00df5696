#!/bin/bash
# Evidence pass for profiles/: launch list of the bench command + full ncu capture of the dominant kernel.
mkdir -p gpurun_out
R=${ROUND:-r1}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
  python bench.py --steps 2 --warmup 1 --no-baselines > gpurun_out/bench_ncu_$R.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel|vx_csr_rows_kernel" -s 4 -c 1 -f -o gpurun_out/prof_${R}_bench \
  python bench.py --steps 2 --warmup 1 --no-baselines > gpurun_out/bench_ncufull_$R.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
