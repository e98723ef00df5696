#!/bin/bash
# R-MAT (C5 shape, reduced scale) at 1 and 4 GPUs on the same box: scaling + balance evidence.
export VX_BENCH_NCCL_TIMEOUT_S=90 VX_BENCH_STACK_DUMP_S=150 TMO=240
SC=${SC:-0.25}
NG=1 WL=rmat25 SCALE=$SC STEPS=10 EXTRA=--no-baselines bash scripts/gpu_multi.sh 2>&1 | cut -c1-700
NG=4 WL=rmat25 SCALE=$SC STEPS=10 bash scripts/gpu_multi.sh 2>&1 | cut -c1-700
