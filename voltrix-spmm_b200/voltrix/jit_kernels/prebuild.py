"""Ahead-of-time population of the JIT cache (every kernel variant the autotuner can pick).

Called by ``__graft_entry__.build()`` on the GPU-less build box: nvcc cross-compiles sm_100a there, the
cache directory is in-tree, so the GPU box starts with a warm cache and never waits on nvcc.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import torch

from . import bmat_swizzle, hmat_gem, preprocess, spmm, spmm_csr, tiles, value_tiles
from .tuner import jit_tuner


def _hmat_arg_defs():
    return hmat_gem.hmat_arg_defs() + (("stream", torch.cuda.Stream),)


def _swizzle_arg_defs():
    return bmat_swizzle.swizzle_arg_defs() + (("stream", torch.cuda.Stream),)


def all_variants():
    """(name, keys, space, includes, arg_defs, template) for every artefact."""
    out = [
        ("preprocess_kernel", {}, tuple(), preprocess.includes, preprocess.arg_defs, preprocess.template),
        ("hmat_gen_kernel", {}, tuple(), hmat_gem.includes, _hmat_arg_defs(), hmat_gem.template),
        ("hmat_packed_swizzle_kernel", {}, tuple(), bmat_swizzle.includes, _swizzle_arg_defs(), bmat_swizzle.template),
        ("csr_tiles_kernel", {}, tuple(), tiles._tiles_includes, tiles._tiles_arg_defs, tiles._tiles_template),
        ("schedule_kernel", {}, tuple(), tiles._sched_includes, tiles._sched_arg_defs, tiles._sched_template),
    ]
    for dtype, ctype in spmm._CTYPE.items():
        space = spmm.SPACE_FP32 if dtype == torch.float32 else \
            spmm.SPACE_HALF + spmm.EXTRA_HALF + tuple(c for c in spmm.SPACE_HALF_NARROW if c.get("ft") == 64)
        out.append(("spmm_kernel", {"ctype": ctype, "weighted": "false", "ft": 128}, space, spmm.includes, spmm.arg_defs_for(dtype),
                    spmm.template))
        wspace = spmm.SPACE_FP32_WEIGHTED if dtype == torch.float32 else spmm.SPACE_HALF_WEIGHTED + \
            ({"model": 0, "stages": 16, "npw": 4},) + tuple(c for c in spmm.SPACE_HALF_NARROW if c.get("ft") == 64)
        if dtype == torch.float32:
            space = space + ({"model": 3, "stages": 24, "npw": 8},) if {"model": 3, "stages": 24, "npw": 8} not in space else space
        out.append(("spmm_kernel", {"ctype": ctype, "weighted": "true", "ft": 128}, wspace, spmm.includes, spmm.arg_defs_for(dtype),
                    spmm.template))
        if dtype != torch.float32:
            out.append(("value_tiles_kernel", {"ctype": ctype}, tuple(), value_tiles.includes, value_tiles.arg_defs_for(dtype),
                        value_tiles.template))
        out.append(("spmm_csr_weighted_kernel", {"ctype": ctype}, tuple(), spmm_csr.includes, spmm_csr.arg_defs_for(dtype),
                    spmm_csr.template))
    return out


def precompile_all(verbose: bool = False):
    variants = all_variants()
    jobs = []
    for name, keys, space, includes, arg_defs, template in variants:
        for code, tuned in jit_tuner.candidates(keys, space, includes, arg_defs, template):
            jobs.append((name, arg_defs, code, tuned))
    from .tuner import _build_one
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda j: _build_one(*j), jobs))
    failed = [(j[0], j[3], r[2]) for j, r in zip(jobs, results) if r[0] is None]
    if failed:
        raise RuntimeError("JIT prebuild failed:\n" + "\n".join(f"{n} {k}: {e}" for n, k, e in failed))
    if verbose:
        for j, r in zip(jobs, results):
            print(f"  built {j[0]} {j[3]} -> {os.path.basename(r[0].path)}")
    return [r[0] for r in results]
