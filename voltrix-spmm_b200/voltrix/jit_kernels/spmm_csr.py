"""``spmm_csr_weighted_kernel``: general CSR x dense with fp32 edge values, CUDA-core rows.

No reference counterpart: the reference's format is binary (every stored entry is 1, bmat_kernels.cuh:102-103) and SURVEY.md
section 8f rank 2 lists weighted A as the next thing real callers (GCN / GraphSAGE) need.  The tile format of the tensor-core path
carries no value array, so weighted products run on the vectorised row kernels (``launch_csr_rows_weighted``); symmetric
normalisation ``D^-1/2 A D^-1/2`` does not need this path at all (``voltrix.spmm_gcn``: scaling folded into the operand and
the epilogue of the tensor-core kernel).
"""
import torch

from ._common import check, current_stream
from .tuner import jit_tuner

includes = ('"voltrix/spmm_kernels.cuh"',)
template = """
voltrix::Epilogue epi;
epi.row_scale = row_scale;
epi.bias = bias;
epi.relu = relu;
__return_code = voltrix::launch_csr_rows_weighted<{ctype}>(indptr, indices, values, num_rows, num_edges, embedding_dim,
                                                           input, output, stream, epi);
"""
_CTYPE = {torch.float32: "float", torch.float16: "__half", torch.bfloat16: "__nv_bfloat16"}


def arg_defs_for(dtype):
    return (("indptr", torch.int32), ("indices", torch.int32), ("values", torch.float32), ("num_rows", int),
            ("num_edges", int), ("embedding_dim", int), ("input", dtype), ("output", torch.float32),
            ("row_scale", torch.float32), ("bias", torch.float32), ("relu", int), ("stream", torch.cuda.Stream))


def spmm_csr_weighted_kernel(indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, num_rows: int,
                             embedding_dim: int, input: torch.Tensor, output: torch.Tensor, row_scale=None, bias=None,
                             relu: bool = False):
    assert indptr.is_cuda and indptr.dtype == torch.int32 and indptr.numel() == num_rows + 1
    assert indices.is_cuda and indices.dtype == torch.int32
    assert values.is_cuda and values.dtype == torch.float32 and values.numel() == indices.numel() and values.is_contiguous()
    assert input.is_cuda and input.dtype in _CTYPE and input.is_contiguous() and input.shape[-1] == embedding_dim
    assert output.is_cuda and output.dtype == torch.float32 and output.is_contiguous()
    assert tuple(output.shape) == (num_rows, embedding_dim)
    if row_scale is not None:
        assert row_scale.is_cuda and row_scale.dtype == torch.float32 and row_scale.numel() == num_rows
    if bias is not None:
        assert bias.is_cuda and bias.dtype == torch.float32 and bias.numel() == embedding_dim
    args = (indptr, indices, values, num_rows, int(indices.numel()), embedding_dim, input, output, row_scale, bias,
            int(bool(relu)), current_stream())
    runtime = jit_tuner.compile_and_tune(name="spmm_csr_weighted_kernel", keys={"ctype": _CTYPE[input.dtype]}, space=tuple(),
                                         includes=includes, arg_defs=arg_defs_for(input.dtype), template=template, args=args)
    check(runtime(*args), "spmm_csr_weighted_kernel")
