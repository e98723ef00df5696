#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2h_t_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/r2h_t_gpu.log | grep -v "Warning\|sparse_csr\|^$"
echo "== suite (low-degree graphs, predicated rounds)"; timeout -s KILL 900 python scripts/suite.py --datasets Yeast YeastH DD amazon0505 com-amazon web-BerkStan ppi protein reddit --out $O/r2h_suite_lowdeg.csv > $O/r2h_suite_lowdeg.log 2>&1; echo "rc=$?"; cat $O/r2h_suite_lowdeg.csv | cut -d, -f1,5,8,9,10,11,14
echo "== shard cost probe"; timeout -s KILL 600 python scripts/shard_cost_probe.py > $O/r2h_shard_cost.log 2>&1; echo "rc=$?"; grep "shard 1/" $O/r2h_shard_cost.log
