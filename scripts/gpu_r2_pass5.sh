#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2e_t_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/r2e_t_gpu.log | grep -v "Warning\|sparse_csr\|^$"
echo "== kineto probe"; timeout -s KILL 600 python scripts/kineto_probe.py > $O/r2e_kineto_probe.log 2>&1; echo "rc=$?"; grep -v Warning $O/r2e_kineto_probe.log | tail -30
echo "== reorder probe"; timeout -s KILL 600 python scripts/reorder_probe.py > $O/r2e_reorder_probe.log 2>&1; echo "rc=$?"; tail -6 $O/r2e_reorder_probe.log
