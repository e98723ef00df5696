"""Aggregate host<->device bandwidth of the box, as seen by the end-to-end leg of bench.py: every rank copies pinned host
buffers to / from its GPU concurrently; k = 1, 2, 4, 8 ranks active.  With and without binding the process (and its pinned
pages, first touch) to the NUMA node of its GPU.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
from voltrix.distributed import bind_to_gpu_numa_node  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
MB = 64


def run(tag):
    h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
    h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
    d_in, d_out = torch.empty(MB << 20, dtype=torch.uint8, device=dev), torch.ones(MB << 20, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("d2h", "h2d", "both"):
        for k in (1, 2, 4, 8):
            if k > world:
                continue
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if rank < k:
                for _ in range(20):
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s1):
                            h_out.copy_(d_out, non_blocking=True)
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s2):
                            d_in.copy_(h_in, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            b.record(); torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                per_dir = 20 * MB * k / 1024 / (ms.item() / 1e3)
                print(f"{tag:10s} {mode:5s} {k} rank(s): {per_dir:7.1f} GiB/s aggregate per direction "
                      f"({per_dir / k:5.1f} per GPU)", flush=True)


run("unbound")
node = bind_to_gpu_numa_node(lr)
if rank == 0:
    print(f"rank 0 bound to NUMA node {node}", flush=True)
run("numa-bound")
dist.destroy_process_group()
