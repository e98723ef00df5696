#!/bin/bash
# Round-1 re-entry GPU pass: tests, bench (both arms), isolation, ncu evidence, other workloads.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tests.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_head.json
echo "== isolate"; timeout 600 python scripts/isolate.py > gpurun_out/isolate.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/isolate.log
echo "== evidence"; ROUND=r1b bash scripts/gpu_evidence.sh
echo "== workloads"; bash scripts/gpu_wl.sh
echo "== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
