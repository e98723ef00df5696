"""``hmat_gen_kernel``: edges -> fp32 16x8 tiles (`hspa`) + column index lists (`hind`).

Reference: voltrix/jit_kernels/hmat_gem.py:13-73 -> voltrix::hmat_cuda (bmat_kernels.cuh:195-212).
Same keyword names; writes exactly the first ``pointer1[-1]`` blocks of the (possibly over-allocated)
``hspa`` / ``hind`` buffers, as the reference does.  One thread per edge instead of one rescan of the
window per TC block; runs on the current stream.
"""
import torch

from ._common import expect_cuda, launch_untuned

includes = ('"voltrix/bmat_kernels.cuh"',)
template = """
__return_code = voltrix::hmat_cuda(node_pointer, edge_list, block_partition, edge_to_column, edge_to_row, pointer1,
                                   num_row_windows, num_nodes, (int64_t)num_edges, hspa, hind, stream);
"""
_I32, _F32 = torch.int32, torch.float32


def hmat_arg_defs():
    """``launch`` signature without the trailing stream (used by the prebuild step, which has no operands)."""
    ints = ("node_pointer", "edge_list", "block_partition", "edge_to_column", "edge_to_row", "pointer1")
    return tuple((n, _I32) for n in ints) + (("num_row_windows", int), ("num_nodes", int), ("num_edges", int),
                                             ("hspa", _F32), ("hind", _I32))


def hmat_gen_kernel(node_pointer: torch.Tensor, edge_list: torch.Tensor, block_partition: torch.Tensor,
                    edge_to_column: torch.Tensor, edge_to_row: torch.Tensor, pointer1: torch.Tensor,
                    hspa: torch.Tensor, hind: torch.Tensor):
    tensors = dict(node_pointer=node_pointer, edge_list=edge_list, block_partition=block_partition,
                   edge_to_column=edge_to_column, edge_to_row=edge_to_row, pointer1=pointer1, hspa=hspa, hind=hind)
    sizes = dict(num_row_windows=int(block_partition.shape[0]), num_nodes=int(node_pointer.shape[0]) - 1,
                 num_edges=int(edge_list.shape[0]))
    defs = hmat_arg_defs()
    expect_cuda(**{n: (tensors[n], t) for n, t in defs if n in tensors})
    launch_untuned("hmat_gen_kernel", includes, template,
                   [(n, t, tensors[n] if n in tensors else sizes[n]) for n, t in defs])
