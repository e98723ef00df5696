#!/bin/bash
O=gpurun_out; mkdir -p $O
V=0/14/7/64,0/28/7/64,0/21/7/64,0/22/11/64,0/33/11/64,0/42/14/64,0/30/10/64,0/20/10/64,0/14/7,0/22/11,1/32/8
for wl in "reddit 1.0 64" "reddit 1.0 32" "FraudYelp-RSR 1.0 64" "amazon0505 1.0 64" "ppi 1.0 64" "ddi 1.0 32" "web-BerkStan 1.0 32"; do
  set -- $wl
  echo "-- $1 N=$3"; timeout -s KILL 400 python scripts/time_models.py --workload $1 --scale $2 --N $3 --only $V 2>&1 | grep "^model\|Error\|error" | cut -c1-84
done | tee $O/r2q_ft64_variants.log
