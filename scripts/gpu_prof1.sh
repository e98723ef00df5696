#!/bin/bash
mkdir -p gpurun_out
for v in 1/16 0/16; do
  tag=${v%%/*}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vx_csr_rows_kernel|vx_spmm_tc_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r1_model$tag \
    python scripts/time_models.py --workload reddit --only $v --once > gpurun_out/ncu_model$tag.log 2>&1
  echo "model $tag rc=$?"; tail -3 gpurun_out/ncu_model$tag.log
done
