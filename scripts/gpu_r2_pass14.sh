#!/bin/bash
O=gpurun_out; mkdir -p $O
V=0/42/14,0/22/11,0/20/10,0/14/7,0/21/7,0/12/6,0/10/5,0/15/5,0/8/4,0/12/4,0/28/14,0/24/12,0/18/6
for wl in "reddit 1.0 128" "rmat25 0.25 256" "amazon0505 1.0 128" "FraudYelp-RSR 1.0 256" "web-BerkStan 1.0 128" "products 1.0 256"; do
  set -- $wl
  echo "-- $1 N=$3"; timeout -s KILL 400 python scripts/time_models.py --workload $1 --scale $2 --N $3 --only $V 2>&1 | grep "^model" | cut -c1-72
done | tee $O/r2n_multi_cta_variants.log
