#!/bin/bash
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/r2_topo.txt 2>&1
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe.py > $O/r2_pcie_probe.log 2>&1; echo "rc=$?"
grep "GiB/s\|bound" $O/r2_pcie_probe.log
