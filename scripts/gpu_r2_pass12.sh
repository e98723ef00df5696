#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== two CTAs per SM probe (reddit, N=128 fp16)"
for v in 0/20/10 0/16/8 0/18/6; do
  timeout -s KILL 300 python scripts/isolate.py --workload reddit --variant $v --flags="|-DVX_TC_CTAS_PER_SM=2" 2>&1 | grep "flags"
done | tee $O/r2l_two_ctas_probe.log
timeout -s KILL 300 python scripts/isolate.py --workload reddit --variant 0/42/14 --flags="" 2>&1 | grep flags | tee -a $O/r2l_two_ctas_probe.log
