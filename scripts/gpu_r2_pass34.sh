#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2ai_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2ai_t_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
