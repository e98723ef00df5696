"""Prebuild specific spmm_kernel variants into the in-tree JIT cache under the CURRENT environment (so that
VOLTRIX_EXTRA_NVCC_FLAGS, which is part of the cache key, applies).  Runs on the GPU-less build box.
    VOLTRIX_EXTRA_NVCC_FLAGS=-DVX_TC_DBG=1 python scripts/prebuild_variants.py 0/36/12 [fp16]"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))
from voltrix.jit_kernels import spmm  # noqa: E402
from voltrix.jit_kernels.tuner import jit_tuner  # noqa: E402

variants = [tuple(int(x) for x in v.split("/")) for v in sys.argv[1].split(",")]
dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
space = tuple({"model": v[0], "stages": v[1], "npw": v[2], "ft": v[3] if len(v) > 3 else 128} for v in variants)
rts = jit_tuner.precompile("spmm_kernel", {"ctype": spmm._CTYPE[dt], "weighted": "false", "ft": 128}, space, spmm.includes, spmm.arg_defs_for(dt),
                           spmm.template)
for v, r in zip(variants, rts):
    print(v, os.path.basename(r.path))
