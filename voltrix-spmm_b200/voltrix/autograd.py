"""Transpose SpMM and the autograd wrapper GNN training needs ("a backward / transpose SpMM").

The reference targets the forward aggregation of GCN / GraphSAGE (voltrix/include/voltrix/bmat_kernels.cuh:16-20) and has
no backward.  ``d(A @ X) / dX`` applied to an upstream gradient G is ``A^T @ G``: the same kernel on the tiles of ``A^T``.
``csr_transpose`` builds the CSR of ``A^T`` on the GPU (one sort by column), ``SparseAdj`` keeps both tile sets (the transpose
lazily, on the first backward) and ``SparseAdj.matmul`` / ``spmm_autograd`` is differentiable w.r.t. the dense operand.
Edge values are treated as constants.
"""
from typing import Optional, Tuple

import torch

from .jit_kernels._common import require_cuda
from .spmm import csr_preprocess, edge_weights as _edge_weights, spmm


def csr_transpose(indptr: torch.Tensor, indices: torch.Tensor, num_cols: Optional[int] = None,
                  values: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, ...]:
    """CSR of ``A^T`` for a CSR matrix ``A`` ``[num_rows, num_cols]``: ``(indptr_t int32 [num_cols + 1], indices_t int32 [nnz]
    [, values_t])``, columns of every row ascending (a stable sort by column keeps the row order).  Runs on the tensors'
    device (torch ops; data plumbing around the hot path, like the graph generators)."""
    num_rows = indptr.numel() - 1
    nnz = indices.numel()
    if num_cols is None:
        num_cols = max(num_rows, int(indices.max().item()) + 1 if nnz else 0)
    dev = indices.device
    deg = (indptr[1:] - indptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(num_rows, device=dev, dtype=torch.int64), deg)
    cols = indices.to(torch.int64)
    order = torch.argsort(cols * num_rows + rows) if nnz else torch.empty(0, dtype=torch.int64, device=dev)
    indices_t = rows[order].to(torch.int32)
    counts = torch.bincount(cols, minlength=num_cols)
    indptr_t = torch.zeros(num_cols + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr_t[1:])
    out = (indptr_t.to(torch.int32), indices_t)
    if values is not None:
        out = out + (values[order],)
    return out


class _SpMMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat: torch.Tensor, adj: "SparseAdj"):
        ctx.adj = adj
        ctx.in_dtype = feat.dtype
        return adj._forward(feat)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        adj = ctx.adj
        g = grad_out.contiguous().to(ctx.in_dtype)        # the transpose product runs in the operand's precision
        return adj._backward(g).to(ctx.in_dtype), None


class SparseAdj:
    """A sparse matrix ``A [num_rows, num_cols]`` preprocessed for ``A @ X`` and, on demand, for ``A^T @ G``.

    ``SparseAdj(indptr, indices, num_rows, num_cols=None, values=None)``; ``adj.matmul(X)`` (or ``adj @ X``) returns the fp32
    product and records the graph, so ``loss.backward()`` reaches ``X.grad`` through the transpose SpMM."""

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, num_rows: int, num_cols: Optional[int] = None,
                 values: Optional[torch.Tensor] = None):
        require_cuda()
        dev = indptr.device if indptr.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.indptr, self.indices = indptr.to(dev).contiguous(), indices.to(dev).contiguous()
        self.values = values.to(dev, torch.float32).contiguous() if values is not None else None
        self.num_rows = int(num_rows)
        self.num_cols = int(num_cols) if num_cols is not None else self.num_rows
        self.nnz = int(self.indices.numel())
        self.tiles = csr_preprocess(self.indptr, self.indices, self.num_rows, num_cols=self.num_cols)
        self.weights = _edge_weights(*self.tiles, self.indptr, self.indices, self.values) if values is not None else None
        self._t = None     # (tiles_t, weights_t), built on the first backward

    def _forward(self, feat: torch.Tensor) -> torch.Tensor:
        assert feat.shape[0] == self.num_cols, "dense operand must have one row per column of A"
        return spmm(*self.tiles, self.num_rows, self.nnz, feat.contiguous(), edge_weights=self.weights)

    def transposed(self):
        if self._t is None:
            t = csr_transpose(self.indptr, self.indices, self.num_cols, self.values)
            tiles_t = csr_preprocess(t[0], t[1], self.num_cols, num_cols=self.num_rows)
            w_t = _edge_weights(*tiles_t, t[0], t[1], t[2]) if self.values is not None else None
            self._t = (tiles_t, w_t)
        return self._t

    def _backward(self, grad: torch.Tensor) -> torch.Tensor:
        tiles_t, w_t = self.transposed()
        return spmm(*tiles_t, self.num_cols, self.nnz, grad, edge_weights=w_t)

    def matmul(self, feat: torch.Tensor) -> torch.Tensor:
        if feat.requires_grad and torch.is_grad_enabled():
            return _SpMMFunction.apply(feat, self)
        return self._forward(feat)

    __matmul__ = matmul


def spmm_autograd(adj: SparseAdj, feat: torch.Tensor) -> torch.Tensor:
    """``adj @ feat`` with gradient w.r.t. ``feat`` (fp32 / fp16 / bf16)."""
    return adj.matmul(feat)
