#!/bin/bash
O=gpurun_out; mkdir -p $O
V=0/42/14,0/20/10,0/22/11,0/21/7,0/14/7,0/12/6,0/18/9,0/16/8,0/24/8
for wl in "reddit 1.0 128" "reddit 1.0 256" "products 1.0 256" "rmat25 0.25 256" "amazon0505 1.0 128" "ppi 1.0 128" "FraudYelp-RSR 1.0 128" "ddi 1.0 128"; do
  set -- $wl
  echo "-- $1 N=$3"; timeout -s KILL 400 python scripts/time_models.py --workload $1 --scale $2 --N $3 --only $V 2>&1 | grep "^model" | cut -c1-100
done | tee $O/r2m_multi_cta_variants.log
