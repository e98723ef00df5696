#!/bin/bash
O=gpurun_out; mkdir -p $O
for wl in Yeast amazon0505; do
  echo "== ncu full $wl model 1"; timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"vx_spmm_csr" -s 1 -c 1 -f -o $O/r2g_prof_csr_$wl \
    python scripts/time_models.py --workload $wl --only 1/32/8 --once > $O/r2g_ncu_$wl.log 2>&1; echo "rc=$?"
done
echo "== variants"; for wl in Yeast amazon0505; do timeout -s KILL 300 python scripts/time_models.py --workload $wl --only 1/32/8,0/42/14,2/32/8 2>&1 | grep -v Warn | tail -4; done
echo "== shard cost probe"; timeout -s KILL 600 python scripts/shard_cost_probe.py > $O/r2g_shard_cost.log 2>&1; echo "rc=$?"; grep "shard 1/" $O/r2g_shard_cost.log
