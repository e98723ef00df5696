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
from ..project import FP32_EXACT_FLAG, FP32_MODE_FLAG
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
plan.sparse_mean_degree = sparse_mean_degree;
plan.input_rows = input_rows;
plan.split_ws = split_ws;
plan.epilogue.row_scale = row_scale;
plan.epilogue.bias = bias;
plan.epilogue.relu = relu;
plan.ticket = ticket;
plan.value_tiles = value_tiles;
plan.csr_values = csr_values;
__return_code = voltrix::voltrix_spmm_forward_cuda<{ctype}, {stages}, {npw}, {weighted}, {ft}>(
    blk_offsets, hspa_packed, hind,
    num_nodes, num_edges, embedding_dim, input, output, {model}, plan, stream);
"""

_CTYPE = {torch.float32: "float", torch.float16: "__half", torch.bfloat16: "__nv_bfloat16"}

# autotune space per input dtype: (model, K-steps in flight per CTA, producer warps per CTA).  The tensor-core variants differ
# in how many CTAs share an SM (csrc/voltrix/spmm_tcgen05.cuh::tc_ctas_per_sm): 14/7 three, 22/11 and 15/5 ... two / three, 42/14
# one.  Several small rings per SM = several MMA-issuing warps; they win on every shape measured (profiles/r2n_multi_cta_variants.txt).
SPACE_HALF = ({"model": 0, "stages": 14, "npw": 7}, {"model": 0, "stages": 22, "npw": 11}, {"model": 0, "stages": 15, "npw": 5},
              {"model": 0, "stages": 42, "npw": 14}, {"model": 1, "stages": 32, "npw": 8}, {"model": 2, "stages": 32, "npw": 8})
# dense operands of at most 64 columns: the 64-wide feature tile (MMA M = 64, one swizzle atom per gathered row), deeper rings
# for the same bytes in flight; the 128-wide 14/7 stays in as the control
SPACE_HALF_NARROW = ({"model": 0, "stages": 20, "npw": 10, "ft": 64}, {"model": 0, "stages": 21, "npw": 7, "ft": 64},
                     {"model": 0, "stages": 33, "npw": 11, "ft": 64}, {"model": 0, "stages": 14, "npw": 7},
                     {"model": 1, "stages": 32, "npw": 8}, {"model": 2, "stages": 32, "npw": 8})
# variants reachable only through the explicit model=/stages=/npw= arguments (tests, scripts): prebuilt as well
EXTRA_HALF = ({"model": 0, "stages": 16, "npw": 4}, {"model": 0, "stages": 40, "npw": 24})
# fp32: model 4 = tcgen05 on ONE fp16 term when the operand is inside fp16's normal range (else model 3's pipeline, decided on
# the device); model 3 = tcgen05 on two bf16 terms (hi + lo); models 1 / 2 = exact-fp32 CUDA-core rows
SPACE_FP32 = ({"model": 4, "stages": 12, "npw": 6}, {"model": 3, "stages": 12, "npw": 6}, {"model": 3, "stages": 24, "npw": 8},
              {"model": 1, "stages": 32, "npw": 8}, {"model": 2, "stages": 32, "npw": 8})
# producer warps of a variant named by its K-step count alone (explicit model=/stages= calls without npw=)
DEFAULT_NPW = {8: 4, 10: 5, 12: 6, 14: 7, 15: 5, 16: 4, 20: 10, 21: 7, 22: 11, 24: 8, 32: 8, 33: 11, 36: 12, 40: 24, 42: 14}

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
        ("sparse_mean_degree", float),
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


def fp32_mode() -> str:
    import os
    return os.environ.get(FP32_MODE_FLAG, "exact" if os.environ.get(FP32_EXACT_FLAG, "0") == "1" else "tf32")


def fp32_space():
    """Autotune space for fp32 input.  The reference's only arithmetic for fp32 operands is TF32 (10 mantissa bits,
    spmm_kernels.cuh:1631-1678); here ``VOLTRIX_FP32_MODE`` picks the precision class, so that fp32 numerics follow a stated
    policy rather than whichever candidate happened to be fastest on this machine:
      ``tf32``  (default) every candidate: model 4 (one fp16 term, 11 bits, when the operand is in fp16's normal range --
                else model 3's pipeline), model 3 (two bf16 terms, 16 bits), the exact CUDA-core rows; results carry at
                least the reference's precision;
      ``split`` model 3 and the exact rows (>= 16 mantissa bits);
      ``exact`` (or the older ``VOLTRIX_FP32_EXACT=1``) exact-fp32 CUDA-core rows only."""
    mode = fp32_mode()
    if mode == "exact":
        return tuple(c for c in SPACE_FP32 if c["model"] in (1, 2))
    if mode == "split":
        return tuple(c for c in SPACE_FP32 if c["model"] != 4)
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
    ft=None,
):
    """Extensions beyond the reference's signature: ``plan`` / ``model`` / ``stages`` / ``npw`` / ``ft`` (explicit variant), the
    fused epilogue ``output = act(row_scale[:, None] * (A @ input) + bias[None, :])`` with fp32 ``row_scale [num_nodes]``,
    fp32 ``bias [embedding_dim]`` and ``act`` = ReLU when ``relu`` (all optional, applied in the kernel that writes C), and
    ``edge_weights`` (``voltrix.edge_weights(...)``): A carries a value per stored entry instead of 1."""
    # Steady state: the same matrix, width, dtype, stream and variant as an earlier call -- replay its marshalled launch
    # with the operand pointers patched in (see _FastLaunch).
    stream = current_stream()
    sid = int(stream.cuda_stream)
    fast_key = (embedding_dim, input.dtype, sid, model, stages, npw, ft, id(edge_weights) if edge_weights is not None else 0,
                id(plan) if plan is not None else 0, fp32_mode() if input.dtype == torch.float32 else "")
    fast = getattr(hspa_packed, "_vx_fast", None)
    if fast is not None:
        hit = fast.get(fast_key)
        if hit is not None and hit.matches(blk_offsets, hind, num_nodes, num_edges, input, output):
            check(hit.launch(input, output, row_scale, bias, relu), "spmm_kernel")
            return
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
    p = plan.launch_args(embedding_dim, sid) if plan is not None else (None, 0, None, 0, None, None, None, None, 0, -1.0)
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
        stages = int(stages or (12 if int(model) in (3, 4) else 14))
        npw = int(npw or DEFAULT_NPW.get(stages, 8))
        space = ({"model": int(model), "stages": stages, "npw": npw, "ft": int(ft or 128)},)
        keys = {"ctype": _CTYPE[input.dtype], "fixed": f"{model}/{stages}/{npw}/{int(ft or 128)}"}
    elif weighted:
        space = SPACE_FP32_WEIGHTED if input.dtype == torch.float32 else \
            tuple(c for c in (SPACE_HALF_NARROW if embedding_dim <= 64 else SPACE_HALF) if c["model"] in (0, 1))
        keys = {"ctype": _CTYPE[input.dtype], "feature_hash": feature_hash(hspa_packed), "N": embedding_dim, "plan": "pcw"}
    else:
        space = fp32_space() if input.dtype == torch.float32 else (SPACE_HALF_NARROW if embedding_dim <= 64 else SPACE_HALF)
        # the key also says what the plan can do: a winner found with the CSR arrays (model 1) or a work list must not
        # be replayed on a matrix that came without them under the same user-chosen hash_tag
        caps = ("p" if plan is not None else "-") + ("c" if plan is not None and plan.csr_indptr is not None else "-")
        keys = {"ctype": _CTYPE[input.dtype], "feature_hash": feature_hash(hspa_packed), "N": embedding_dim, "plan": caps}

    keys["weighted"] = "true" if weighted else "false"
    if plan is not None and model is None and not getattr(plan, "route_is_default", True):
        # a non-default routing rule (voltrix.reschedule / tune_routing) keeps its own winner even under a shared hash_tag;
        # the default rule and the rule-less plan (no CSR arrays) keep the keys they always had
        keys["route"] = f"{plan.sparse_ratio:g}/{plan.small_blocks}"
    keys["ft"] = 128          # default feature tile; a candidate that names its own overrides it
    if input.dtype == torch.float32 and model is None:
        keys["fp32"] = fp32_mode()      # winners are per precision class

    # fp32 on the tensor cores (model 3) needs a bf16 [rows, 2N] workspace: hand it over while that model is still a
    # candidate for this key, drop it once the tuner has settled on a CUDA-core model
    signature = ("spmm_kernel", f"{ {k: keys[k] for k in sorted(keys)} }")
    winner = jit_tuner.tuned_keys.get(signature)
    ws_owner = plan if plan is not None else hspa_packed
    split_ws = None
    if input.dtype == torch.float32 and embedding_dim % 8 == 0 and any(c["model"] in (3, 4) for c in space) and \
            (winner is None or winner.get("model") in (3, 4)):
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
    if split_ws is not None and jit_tuner.tuned_keys.get(signature, {}).get("model") not in (3, 4):
        getattr(ws_owner, "_vx_split_ws", {}).clear()
        split_ws = None
        args = args[:ARG_SPLIT_WS] + (None,) + args[ARG_SPLIT_WS + 1:]
    try:
        if fast is None:
            fast = hspa_packed._vx_fast = {}
        fast[fast_key] = _FastLaunch(runtime, args, blk_offsets, hind, num_nodes, num_edges, input, output)
    except AttributeError:      # a tensor subclass that refuses attributes: every call takes the full path
        pass


_ARG_NAMES = tuple(n for n, _ in arg_defs_for(torch.float16))
ARG_INPUT, ARG_OUTPUT, ARG_SPLIT_WS = _ARG_NAMES.index("input"), _ARG_NAMES.index("output"), _ARG_NAMES.index("split_ws")
ARG_ROW_SCALE, ARG_BIAS, ARG_RELU = _ARG_NAMES.index("row_scale"), _ARG_NAMES.index("bias"), _ARG_NAMES.index("relu")


class _FastLaunch:
    """One fully resolved SpMM launch (variant chosen, plan / scratch / ticket pointers marshalled): later calls with the
    same matrix, width, dtype and stream patch five slots -- input, output, row_scale, bias, relu -- and go straight to the
    artefact's ``launch``.  The tensors whose pointers are baked in are kept alive here."""

    def __init__(self, runtime, args, blk_offsets, hind, num_nodes, num_edges, input, output):
        import ctypes
        self._c = ctypes
        self.runtime = runtime
        self.keep = args
        self.cvals = runtime.prepare(args)
        self.ids = (blk_offsets.data_ptr(), hind.data_ptr(), num_nodes, num_edges)
        self.in_shape, self.out_shape = tuple(input.shape), tuple(output.shape)

    def matches(self, blk_offsets, hind, num_nodes, num_edges, input, output) -> bool:
        return (self.ids == (blk_offsets.data_ptr(), hind.data_ptr(), num_nodes, num_edges)
                and tuple(input.shape) == self.in_shape and tuple(output.shape) == self.out_shape
                and input.is_cuda and output.is_cuda and output.dtype == torch.float32
                and input.is_contiguous() and output.is_contiguous())

    def launch(self, input, output, row_scale, bias, relu) -> int:
        vp = self._c.c_void_p
        cv = list(self.cvals)      # private copy: two Python threads may launch the same matrix
        cv[ARG_INPUT] = vp(input.data_ptr())
        cv[ARG_OUTPUT] = vp(output.data_ptr())
        if row_scale is not None:
            assert row_scale.is_cuda and row_scale.dtype == torch.float32 and row_scale.numel() == self.out_shape[0]
            cv[ARG_ROW_SCALE] = vp(row_scale.data_ptr())
        else:
            cv[ARG_ROW_SCALE] = vp(None)
        if bias is not None:
            assert bias.is_cuda and bias.dtype == torch.float32 and bias.numel() == self.out_shape[1]
            cv[ARG_BIAS] = vp(bias.data_ptr())
        else:
            cv[ARG_BIAS] = vp(None)
        cv[ARG_RELU] = self._c.c_int(1 if relu else 0)
        return self.runtime.launch_prepared(cv)
