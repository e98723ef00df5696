#!/bin/bash
mkdir -p gpurun_out
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "rc=$?"; cat gpurun_out/bench1.json; tail -20 gpurun_out/bench1.err
echo "== time_models reddit"; timeout 600 python scripts/time_models.py --workload reddit > gpurun_out/tm_reddit.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/tm_reddit.log
echo "== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
