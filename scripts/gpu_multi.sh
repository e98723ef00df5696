#!/bin/bash
# Multi-GPU bench as the driver launches it.
mkdir -p gpurun_out
N=${NG:-2}
nvidia-smi -L > gpurun_out/smi_multi.txt
echo "== reference arm N=$N"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref_n$N.json
echo "== product arm N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-1800 gpurun_out/bench_n$N.json; grep -E "rank|Error|error" gpurun_out/bench_n$N.err | tail -8
