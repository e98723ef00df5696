#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/suite.py --out gpurun_out/suite_r1.csv > gpurun_out/suite.log 2>&1; echo "rc=$?"; tail -70 gpurun_out/suite.log | cut -c1-200
