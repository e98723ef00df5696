#!/bin/bash
# Round-1 verification pass: GPU tests, bench line, bottleneck isolation, products timing, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 700 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
echo "== isolate"; timeout 300 python scripts/isolate.py --variant 0/36/12 > gpurun_out/isolate.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/isolate.log
echo "== reddit variants"; timeout 300 python scripts/time_models.py --workload reddit --only 0/36/12,0/40/16,0/40/24,1/32/8 > gpurun_out/tm_reddit.log 2>&1; echo "rc=$?"; grep -E "^model|M=" gpurun_out/tm_reddit.log
echo "== products"; timeout 400 python scripts/time_models.py --workload products --iters 5 --only 0/36/12,0/40/16,1/32/8 > gpurun_out/tm_products.log 2>&1; echo "rc=$?"; grep -E "^model|M=" gpurun_out/tm_products.log
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_tc_kernel" -s 1 -c 1 -f -o gpurun_out/prof_tc_r1b \
    python scripts/time_models.py --workload reddit --only 0/36/12 --once > gpurun_out/ncu_tc.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tc.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 2 --warmup 1 --no-baselines > gpurun_out/bench_ncu.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -15
