#!/bin/bash
# Round 2 multi-GPU pass: NG ranks on one box.  The 2-rank NCCL test, then bench.py exactly as the driver launches it
# (headline Reddit-shaped C2 + the products-shaped C4 and R-MAT C5 extras at this N).
NG=${NG:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$NG" = "2" ]; then
  echo "== 2-rank NCCL test"; timeout -s KILL 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 500 > $O/r2z_t_multigpu.log 2>&1; echo "rc=$?"; tail -5 $O/r2z_t_multigpu.log
fi
echo "== bench --gpus $NG"
VX_BENCH_NCCL_TIMEOUT_S=300 timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $NG --steps 20 --warmup 5 > $O/r2z_bench_n$NG.json 2> $O/r2z_bench_n$NG.err; echo "rc=$?"
cut -c1-400 $O/r2z_bench_n$NG.json; grep -v Warning $O/r2z_bench_n$NG.err | grep "rank 0\|rror\|Traceback" | tail -20
