#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== competitors incl. DTC-SpMM (bench_all: ddi, amazon0505, FraudYelp-RSR, reddit; N = 128, 256, 512)"
cd bench && timeout -s KILL 2400 python bench_all.py --datasets ddi amazon0505 FraudYelp-RSR --feature_dims 128 256 512 --results ../$O/r2s_competitors.csv > ../$O/r2s_competitors.log 2>&1; echo "rc=$?"; cd ..
cat $O/r2s_competitors.csv
grep -i "fail\|error" $O/r2s_competitors.log | head
