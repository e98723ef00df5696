#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/c5_breakdown_probe.py rmat25 1.0 2>&1 | grep -v Warning | tail -14 | tee gpurun_out/r2z_c5_breakdown.txt
timeout -s KILL 600 python scripts/c5_breakdown_probe.py products 1.0 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r2z_c4_breakdown.txt
