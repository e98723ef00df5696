#!/bin/bash
# pass 20: competitor sweep with DTC-SpMM (relinked against the shared libstdc++), incl. the reordered DTC column
mkdir -p gpurun_out && rm -f gpurun_out/r2u_competitors.csv
cd bench
timeout -s KILL 1500 python bench_all.py --datasets ddi amazon0505 --feature_dims 128 256 512 --reorder \
   --results ../gpurun_out/r2u_competitors.csv > ../gpurun_out/r2u_competitors.log 2>&1; echo "rc=$?"
grep -c . ../gpurun_out/r2u_competitors.csv; grep "DTC" ../gpurun_out/r2u_competitors.log | head -20; tail -5 ../gpurun_out/r2u_competitors.log
