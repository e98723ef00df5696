#!/bin/bash
mkdir -p gpurun_out
for wl in products c1; do
  timeout 600 python scripts/time_models.py --workload $wl --iters 5 --only 0/36/12,0/32/8,1/32/8 > gpurun_out/tm_$wl.log 2>&1; echo "$wl rc=$?"; grep -E "^model|M=" gpurun_out/tm_$wl.log | tail -5
done
timeout 600 python scripts/time_models.py --workload rmat25 --scale 0.03125 --iters 5 --only 0/36/12,1/32/8 > gpurun_out/tm_rmat20.log 2>&1; echo "rmat rc=$?"; grep -E "^model|M=" gpurun_out/tm_rmat20.log | tail -5
