"""``preprocess_kernel``: the kernel-level column-compaction step.

Reference: voltrix/jit_kernels/preprocess.py:23-70 -> voltrix::preprocess (host C++, single thread,
bmat_kernels.cuh:264-320).  Same keyword names and the same four outputs (block_partition,
edge_to_column, edge_to_row, pointer1), but computed by GPU kernels (csrc/voltrix/bmat_kernels.cuh).
The reference only takes CPU tensors; here CPU tensors are staged through the device and the results
copied back into the caller's tensors, CUDA tensors are used in place.
"""
import torch

from ._common import alloc_workspace, check, current_stream, require_cuda, ws_units
from .tuner import jit_tuner

includes = ('"voltrix/bmat_kernels.cuh"',)
template = """
ws_query[0] = (int64_t)voltrix::preprocess_workspace_bytes(num_edges, num_nodes);
if (workspace == nullptr) { __return_code = 0; return; }
__return_code = voltrix::preprocess(
    edge_list, node_pointer, num_nodes, (int64_t)num_edges, BLK_H, BLK_W,
    block_partition, edge_to_column, edge_to_row, pointer1,
    workspace, (size_t)workspace_units * 256, stream);
"""

arg_defs = (
    ("edge_list", torch.int),
    ("node_pointer", torch.int),
    ("num_nodes", int),
    ("block_partition", torch.int),
    ("edge_to_column", torch.int),
    ("edge_to_row", torch.int),
    ("pointer1", torch.int),
    ("num_edges", int),
    ("workspace", torch.uint8),
    ("workspace_units", int),
    ("ws_query", torch.int64),
    ("stream", torch.cuda.Stream),
)


def _runtime(args):
    return jit_tuner.compile_and_tune(name="preprocess_kernel", keys={}, space=tuple(), includes=includes,
                                      arg_defs=arg_defs, template=template, args=args)


def preprocess_kernel(
    edge_list: torch.Tensor,
    node_pointer: torch.Tensor,
    block_partition: torch.Tensor,
    edge_to_column: torch.Tensor,
    edge_to_row: torch.Tensor,
    pointer1: torch.Tensor,
):
    for t in (edge_list, node_pointer, block_partition, edge_to_column, edge_to_row, pointer1):
        assert t.dtype == torch.int32 and t.is_contiguous()
    require_cuda()
    num_nodes = node_pointer.shape[0] - 1
    num_edges = edge_list.shape[0]
    assert block_partition.numel() >= (num_nodes + 15) // 16 and pointer1.numel() >= block_partition.numel() + 1
    assert edge_to_column.numel() >= num_edges and edge_to_row.numel() >= num_edges

    outs = (block_partition, edge_to_column, edge_to_row, pointer1)
    dev = edge_list.device if edge_list.is_cuda else torch.device("cuda", torch.cuda.current_device())
    d_in = [t if t.is_cuda else t.to(dev) for t in (edge_list, node_pointer)]
    d_out = [t if t.is_cuda else torch.zeros_like(t, device=dev) for t in outs]

    query = torch.zeros(1, dtype=torch.int64)
    stream = current_stream()
    base = (d_in[0], d_in[1], num_nodes, d_out[0], d_out[1], d_out[2], d_out[3], num_edges)
    runtime = _runtime(base + (None, 0, query, stream))
    check(runtime(*base, None, 0, query, stream), "preprocess_kernel (workspace query)")
    ws = alloc_workspace(int(query[0]), dev)
    check(runtime(*base, ws, ws_units(ws.numel()), query, stream), "preprocess_kernel")
    for host, devt in zip(outs, d_out):
        if not host.is_cuda:
            host.copy_(devt)  # synchronises with the stream
