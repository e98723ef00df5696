#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/hybrid_probe.py > gpurun_out/hybrid.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/hybrid.log
