"""``value_tiles_kernel``: per-edge values of a CSR matrix scattered into the tile format's 16 x 8 blocks.

No reference counterpart -- the reference's format is binary (every stored entry is 1, bmat_kernels.cuh:100-103); SURVEY.md
section 8f rank 2 lists weighted A as the next thing GCN / GraphSAGE / attention callers need.  The tiles are what the
WEIGHTED instantiation of the tensor-core SpMM kernel reads in place of the bitmaps (``spmm_kernel(..., edge_weights=)``):
256 bytes of fp16 / bf16 per TC block, laid out as the shared-memory image of the kernel's A^T operand chunk.
"""
import torch

from ._common import check, current_stream
from .tuner import jit_tuner

includes = ('"voltrix/bmat_kernels.cuh"',)
template = """
__return_code = voltrix::value_tiles<{ctype}>(indptr, indices, values, num_nodes, (int64_t)num_edges, blk_offsets, hind,
                                              (int64_t)total_blocks, tiles, not_found, stream);
"""
_CTYPE = {torch.float16: "__half", torch.bfloat16: "__nv_bfloat16"}


def arg_defs_for(dtype):
    return (("indptr", torch.int32), ("indices", torch.int32), ("values", torch.float32), ("num_nodes", int),
            ("num_edges", int), ("blk_offsets", torch.int32), ("hind", torch.int32), ("total_blocks", int),
            ("tiles", dtype), ("not_found", torch.int32), ("stream", torch.cuda.Stream))


def value_tiles_kernel(indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, num_nodes: int,
                       blk_offsets: torch.Tensor, hind: torch.Tensor, tiles: torch.Tensor, not_found: torch.Tensor):
    """``tiles``: fp16 / bf16 ``[total_blocks * 128]`` (zeroed here), ``not_found``: int32 ``[1]`` -- stored entries whose
    column does not occur in their window's ``hind`` list (a triple that was not built from this CSR matrix)."""
    assert indptr.is_cuda and indptr.dtype == torch.int32 and indptr.numel() == num_nodes + 1
    assert indices.is_cuda and indices.dtype == torch.int32
    assert values.is_cuda and values.dtype == torch.float32 and values.numel() == indices.numel() and values.is_contiguous()
    assert blk_offsets.is_cuda and blk_offsets.dtype == torch.int32 and hind.is_cuda and hind.dtype == torch.int32
    assert tiles.is_cuda and tiles.dtype in _CTYPE and tiles.is_contiguous() and tiles.numel() % 128 == 0
    assert not_found.is_cuda and not_found.dtype == torch.int32
    args = (indptr, indices, values, num_nodes, int(indices.numel()), blk_offsets, hind, tiles.numel() // 128, tiles,
            not_found, current_stream())
    runtime = jit_tuner.compile_and_tune(name="value_tiles_kernel", keys={"ctype": _CTYPE[tiles.dtype]}, space=tuple(),
                                         includes=includes, arg_defs=arg_defs_for(tiles.dtype), template=template, args=args)
    check(runtime(*args), "value_tiles_kernel")
