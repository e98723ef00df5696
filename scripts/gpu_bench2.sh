#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/time_models.py --workload reddit --only 0/36/12,0/32/8,1/32/8 > gpurun_out/tm_exp.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tm_exp.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "rc=$?"; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
