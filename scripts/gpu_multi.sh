#!/bin/bash
# Multi-GPU bench as the driver launches it.  NG=<gpus> WL=<workload> SCALE=<scale> STEPS=<k>
mkdir -p gpurun_out
N=${NG:-2}; WL=${WL:-reddit}; SCALE=${SCALE:-1.0}; STEPS=${STEPS:-20}
TAG=${WL}_s${SCALE}_n$N
if [ "$N" = "1" ]; then
  timeout ${TMO:-900} python bench.py --gpus 1 --steps $STEPS --warmup 5 --workload $WL --scale $SCALE ${EXTRA} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
else
  timeout ${TMO:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 5 --workload $WL --scale $SCALE ${EXTRA} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
fi
echo "rc=$?"; cut -c1-900 gpurun_out/bench_$TAG.json; grep -E "^\[rank|Error|error|Traceback" gpurun_out/bench_$TAG.err | tail -12
