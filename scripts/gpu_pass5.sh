#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_gpu.log | cut -c1-250
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
