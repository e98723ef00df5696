"""``hmat_packed_swizzle_kernel``: fp32 16x8 tiles -> 4 x uint32 bitmaps in mma-fragment order.

Reference: voltrix/jit_kernels/bmat_swizzle.py:14-48 -> voltrix::hmat_packed_swizzle_cuda
(bmat_kernels.cuh:228-242, bit order :180-184).  One warp per TC block and one ballot per word here.
"""
import torch

from ._common import expect_cuda, launch_untuned

includes = ('"voltrix/bmat_kernels.cuh"',)
template = """
__return_code = voltrix::hmat_packed_swizzle_cuda(num_row_windows, pointer1, hspa, hspa_packed, stream);
"""


def swizzle_arg_defs():
    return (("num_row_windows", int), ("pointer1", torch.int32), ("hspa", torch.float32), ("hspa_packed", torch.uint32))


def hmat_packed_swizzle_kernel(block_partition: torch.Tensor, pointer1: torch.Tensor, hspa: torch.Tensor,
                               hspa_packed: torch.Tensor):
    expect_cuda(block_partition=(block_partition, torch.int32), pointer1=(pointer1, torch.int32),
                hspa=(hspa, torch.float32), hspa_packed=(hspa_packed, torch.uint32))
    values = dict(num_row_windows=int(block_partition.shape[0]), pointer1=pointer1, hspa=hspa, hspa_packed=hspa_packed)
    launch_untuned("hmat_packed_swizzle_kernel", includes, template, [(n, t, values[n]) for n, t in swizzle_arg_defs()])
