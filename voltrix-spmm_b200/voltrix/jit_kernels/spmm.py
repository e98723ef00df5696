"""``spmm_kernel``: the SpMM launch, JIT-specialised and autotuned.

Reference: voltrix/jit_kernels/spmm.py:17-94.  Same positional / keyword arguments and the same
autotune protocol (``jit_tuner.compile_and_tune`` over a ``model`` space, tuning key derived from
``hspa_packed.hash_tag``), with these extensions:

* ``input`` may be fp32 (exact fp32 CUDA-core paths), fp16 or bf16 (tcgen05 path, fp32 accumulate);
* the tuning key also carries N and the dtype;
* an optional ``plan`` (hung off ``hspa_packed`` by ``csr_preprocess``) supplies the nnz-balanced work
  list and the CSR arrays; without it only the reference triple is used;
* the generated ``launch`` writes ``__return_code``; candidates that cannot run a configuration
  report a non-zero code and are skipped by the tuner instead of killing the process.
"""
import warnings

import torch

from ..jit.compiler import hash_to_hex
from ._common import check, current_stream
from .tuner import jit_tuner

includes = ('"voltrix/spmm_kernels.cuh"',)
template = """
voltrix::SpmmPlan plan;
plan.items = reinterpret_cast<const voltrix::WorkItem*>(items);
plan.num_items = num_items;
plan.fixups = reinterpret_cast<const voltrix::FixupItem*>(fixups);
plan.num_fixups = num_fixups;
plan.scratch = scratch;
plan.csr_indptr = csr_indptr;
plan.csr_indices = csr_indices;
plan.sparse_rows = sparse_rows;
plan.num_sparse_rows = num_sparse_rows;
plan.input_rows = input_rows;
plan.split_ws = split_ws;
plan.epilogue.row_scale = row_scale;
plan.epilogue.bias = bias;
plan.epilogue.relu = relu;
plan.ticket = ticket;
plan.value_tiles = value_tiles;
plan.csr_values = csr_values;
__return_code = voltrix::voltrix_spmm_forward_cuda<{ctype}, {stages}, {npw}, {weighted}>(
    blk_offsets, hspa_packed, hind,
    num_nodes, num_edges, embedding_dim, input, output, {model}, plan, stream);
"""

_CTYPE = {torch.float32: "float", torch.float16: "__half", torch.bfloat16: "__nv_bfloat16"}

# autotune space per input dtype: (model, stages)
SPACE_HALF = ({"model": 0, "stages": 36, "npw": 12}, {"model": 0, "stages": 42, "npw": 14}, {"model": 0, "stages": 32, "npw": 16},
              {"model": 0, "stages": 40, "npw": 24}, {"model": 1, "stages": 32, "npw": 8}, {"model": 2, "stages": 32, "npw": 8})
# variants reachable only through the explicit model=/stages= arguments (tests, scripts): prebuilt as well
EXTRA_HALF = ({"model": 0, "stages": 16, "npw": 4}, {"model": 0, "stages": 32, "npw": 8})
# fp32: model 3 = tcgen05 on two bf16 terms (hi + lo); models 1 / 2 = exact-fp32 CUDA-core rows
SPACE_FP32 = ({"model": 3, "stages": 24, "npw": 8}, {"model": 1, "stages": 32, "npw": 8}, {"model": 2, "stages": 32, "npw": 8})


# A with per-edge values: the WEIGHTED tensor-core instantiations and the weighted CUDA-core CSR rows
SPACE_HALF_WEIGHTED = tuple(c for c in SPACE_HALF if c["model"] in (0, 1))
SPACE_FP32_WEIGHTED = ({"model": 1, "stages": 32, "npw": 8},)


def arg_defs_for(dtype):
    return (
        ("blk_offsets", torch.int32),
        ("hspa_packed", torch.uint32),
        ("hind", torch.int32),
        ("num_nodes", int),
        ("num_edges", int),
        ("embedding_dim", int),
        ("input", dtype),
        ("output", torch.float32),
        ("items", torch.int32),
        ("num_items", int),
        ("fixups", torch.int32),
        ("num_fixups", int),
        ("scratch", torch.float32),
        ("csr_indptr", torch.int32),
        ("csr_indices", torch.int32),
        ("sparse_rows", torch.int32),
        ("num_sparse_rows", int),
        ("input_rows", int),
        ("split_ws", torch.bfloat16),
        ("row_scale", torch.float32),
        ("bias", torch.float32),
        ("relu", int),
        ("ticket", torch.int32),
        ("value_tiles", dtype if dtype != torch.float32 else torch.float16),
        ("csr_values", torch.float32),
        ("stream", torch.cuda.Stream),
    )


def _owner_cache(owner, attr: str) -> dict:
    cache = getattr(owner, attr, None)
    if cache is None:
        cache = {}
        try:
            setattr(owner, attr, cache)
        except AttributeError:
            pass
    return cache


def _split_workspace(owner, rows: int, embedding_dim: int, device, stream_id: int) -> torch.Tensor:
    """bf16 [rows, 2 * embedding_dim] buffer for the fp32 tensor-core path (model 3), cached on the plan (or on
    ``hspa_packed`` when the caller came with a bare reference triple), one per CUDA stream: two SpMMs on the same
    matrix from different streams must not share it."""
    cache = _owner_cache(owner, "_vx_split_ws")
    key = (rows, embedding_dim, str(device), stream_id)
    buf = cache.get(key)
    if buf is None:
        buf = torch.empty((rows, 2 * embedding_dim), dtype=torch.bfloat16, device=device)
        cache[key] = buf
    return buf


def _ticket(owner, device, stream_id: int) -> torch.Tensor:
    """The 4-byte work-claim counter of the tensor-core kernel (zeroed by every launch, on its stream), one per
    (matrix, stream)."""
    cache = _owner_cache(owner, "_vx_ticket")
    key = (str(device), stream_id)
    buf = cache.get(key)
    if buf is None:
        buf = torch.zeros(4, dtype=torch.int32, device=device)
        cache[key] = buf
    return buf


def fp32_space():
    """Autotune space for fp32 input.  Model 3 (two bf16 terms on the tensor cores, ~16 mantissa bits) is left out when
    ``VOLTRIX_FP32_EXACT=1``: then only the exact-fp32 CUDA-core models compete and fp32 results do not depend on which
    candidate happened to be fastest on this machine."""
    import os
    if os.environ.get("VOLTRIX_FP32_EXACT", "0") == "1":
        return tuple(c for c in SPACE_FP32 if c["model"] != 3)
    return SPACE_FP32


def feature_hash(feature: torch.Tensor) -> str:
    """Tuning key of a matrix (reference: jit_kernels/spmm.py:17-36): ``hash_tag`` if set, else the address."""
    if hasattr(feature, "hash_tag") and isinstance(feature.hash_tag, str):
        return hash_to_hex(feature.hash_tag)
    plan = getattr(feature, "_vx_plan", None)
    if plan is not None:
        return hash_to_hex(plan.signature())   # shape statistics of the matrix: stable across processes
    warnings.warn(
        "The feature tensor(i.e. `hspa_packed`)'s hash_tag attr is not set. "
        "Voltrix will use the memory address as the key value for profiling, "
        "which may lead to performance degradation of different cases."
    )
    return hash_to_hex(str(feature.data_ptr()))


def spmm_kernel(
    blk_offsets: torch.Tensor,  # pointer1
    hspa_packed: torch.Tensor,
    hind: torch.Tensor,
    num_nodes: int,
    num_edges: int,
    embedding_dim: int,
    input: torch.Tensor,
    output: torch.Tensor,
    plan=None,
    model=None,
    stages=None,
    npw=None,
    row_scale=None,
    bias=None,
    relu=False,
    edge_weights=None,
):
    """Extensions beyond the reference's signature: ``plan`` / ``model`` / ``stages`` / ``npw`` (explicit variant), the
    fused epilogue ``output = act(row_scale[:, None] * (A @ input) + bias[None, :])`` with fp32 ``row_scale [num_nodes]``,
    fp32 ``bias [embedding_dim]`` and ``act`` = ReLU when ``relu`` (all optional, applied in the kernel that writes C), and
    ``edge_weights`` (``voltrix.edge_weights(...)``): A carries a value per stored entry instead of 1."""
    assert blk_offsets.is_cuda and blk_offsets.dtype == torch.int32
    assert hspa_packed.is_cuda and hspa_packed.dtype == torch.uint32
    assert hind.is_cuda and hind.dtype == torch.int32
    assert input.is_cuda and input.dtype in _CTYPE and input.is_contiguous()
    assert output.is_cuda and output.dtype == torch.float and output.is_contiguous()
    assert input.shape[-1] == embedding_dim and output.shape[-1] == embedding_dim
    if row_scale is not None:
        assert row_scale.is_cuda and row_scale.dtype == torch.float32 and row_scale.is_contiguous()
        assert row_scale.numel() == num_nodes
    if bias is not None:
        assert bias.is_cuda and bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == embedding_dim

    if plan is None:
        plan = getattr(hspa_packed, "_vx_plan", None)
    stream = current_stream()
    sid = int(stream.cuda_stream)
    p = plan.launch_args(embedding_dim, sid) if plan is not None else (None, 0, None, 0, None, None, None, None, 0)
    weighted = edge_weights is not None
    value_tiles = csr_values = None
    if weighted:
        assert plan is not None and plan.items is not None, "edge_weights need the plan csr_preprocess attaches to hspa_packed"
        csr_values = edge_weights.csr_values
        assert csr_values.is_cuda and csr_values.dtype == torch.float32 and csr_values.numel() == num_edges
        if input.dtype != torch.float32:   # the value tiles must be in the dense operand's 16-bit format
            value_tiles = edge_weights.tiles(input.dtype)
            assert value_tiles.numel() == plan.total_blocks * 128
    if model is not None:   # explicit variant (tests, benchmarks): a space of one, no timing runs
        stages = int(stages or (24 if int(model) == 3 else 32))
        npw = int(npw or {8: 4, 16: 4, 36: 12, 42: 14, 40: 24}.get(stages, 8))
        space = ({"model": int(model), "stages": stages, "npw": npw},)
        keys = {"ctype": _CTYPE[input.dtype], "fixed": f"{model}/{stages}/{npw}"}
    elif weighted:
        space = SPACE_FP32_WEIGHTED if input.dtype == torch.float32 else SPACE_HALF_WEIGHTED
        keys = {"ctype": _CTYPE[input.dtype], "feature_hash": feature_hash(hspa_packed), "N": embedding_dim, "plan": "pcw"}
    else:
        space = fp32_space() if input.dtype == torch.float32 else SPACE_HALF
        # the key also says what the plan can do: a winner found with the CSR arrays (model 1) or a work list must not
        # be replayed on a matrix that came without them under the same user-chosen hash_tag
        caps = ("p" if plan is not None else "-") + ("c" if plan is not None and plan.csr_indptr is not None else "-")
        keys = {"ctype": _CTYPE[input.dtype], "feature_hash": feature_hash(hspa_packed), "N": embedding_dim, "plan": caps}

    keys["weighted"] = "true" if weighted else "false"

    # fp32 on the tensor cores (model 3) needs a bf16 [rows, 2N] workspace: hand it over while that model is still a
    # candidate for this key, drop it once the tuner has settled on a CUDA-core model
    signature = ("spmm_kernel", f"{ {k: keys[k] for k in sorted(keys)} }")
    winner = jit_tuner.tuned_keys.get(signature)
    ws_owner = plan if plan is not None else hspa_packed
    split_ws = None
    if input.dtype == torch.float32 and embedding_dim % 8 == 0 and any(c["model"] == 3 for c in space) and \
            (winner is None or winner.get("model") == 3):
        split_ws = _split_workspace(ws_owner, int(input.shape[0]), embedding_dim, input.device, sid)
    ticket = _ticket(ws_owner, input.device, sid)
    args = (blk_offsets, hspa_packed, hind, num_nodes, num_edges, embedding_dim, input, output, *p,
            int(input.shape[0]), split_ws, row_scale, bias, int(bool(relu)), ticket, value_tiles, csr_values, stream)

    runtime = jit_tuner.compile_and_tune(
        name="spmm_kernel",
        keys=keys,
        space=space,
        includes=includes,
        arg_defs=arg_defs_for(input.dtype),
        template=template,
        args=args,
        kernel_tag="spmm",
    )
    check(runtime(*args), "spmm_kernel")
    if split_ws is not None and jit_tuner.tuned_keys.get(signature, {}).get("model") != 3:
        getattr(ws_owner, "_vx_split_ws", {}).clear()
