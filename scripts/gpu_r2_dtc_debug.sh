#!/bin/bash
mkdir -p /tmp/dtcdbg && cd /tmp/dtcdbg
python $GRAFT_REPO_ROOT/bench/graph_gen.py --data_name ddi --num_feats 128 --mtx_max_nnz 0 > gen.log 2>&1; echo "gen rc=$?"
timeout -s KILL 300 python $GRAFT_REPO_ROOT/bench/bm_dtc.py 2>&1 | tail -25; echo "rc=${PIPESTATUS[0]}"
ls -la | head -20; cat DTCSpMM_exe_time_and_throughput.csv 2>/dev/null | head -3
