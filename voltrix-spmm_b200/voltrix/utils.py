"""Metrics and timing helpers used by the tests, the tuner and the bench scripts.

Same names and return conventions as the reference's voltrix/utils.py (``calc_diff`` :38,
``relative_error`` :21, ``GPU_bench`` :324 -> ms, ``bench_kineto`` :232 -> seconds, ``CPU_bench`` :353,
``DurationTimer`` :146).  ``bench_kineto`` differs in one documented way: the reference asserts that
exactly ONE profiler row contains ``kernel_names`` (:298-303); here all matching rows are summed, because
one SpMM call may launch the tcgen05 kernel plus the CUDA-core row kernel and the fix-up pass.
"""
import time
from typing import Tuple, Union

import torch


def check_nan_inf(x: torch.Tensor):
    n, i = torch.isnan(x).sum().item(), torch.isinf(x).sum().item()
    if n or i:
        import warnings
        warnings.warn(f"with {n} nans and {i} infs")


def relative_error(value: torch.Tensor, real: torch.Tensor, exclude_zeros=True) -> float:
    """Mean |value - real| / |real| in fp64 (reference utils.py:21-35)."""
    value = value.double().flatten()
    real = real.double().flatten()
    if not exclude_zeros:
        return ((value - real).abs() / (real.abs() + 1e-9)).mean().item()
    mask = (real.abs() == 0) | real.isinf() | value.isinf()
    value, real = value[~mask], real[~mask]
    return ((value - real).abs() / real.abs()).mean().item()


def calc_diff(x, y, dtype=torch.float):
    """The reference's 'difference rate': 1 - 2<x,y> / (|x|^2 + |y|^2) (utils.py:38-42)."""
    x, y = x.to(dtype), y.to(dtype)
    denominator = (x * x + y * y).sum()
    sim = 2 * (x * y).sum() / denominator
    return 1 - sim


class DurationTimer:
    """CUDA-event (or wall-clock) timer; ``get_duration()`` in ms."""

    def __init__(self, is_sync: bool = True, cuda: bool = True):
        self.cuda = cuda and torch.cuda.is_available()
        self.is_sync = is_sync
        self.duration = 0.0

    def __enter__(self):
        if self.cuda:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            if self.is_sync:
                torch.cuda.synchronize()
            self.start.record()
        else:
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if self.cuda:
            self.end.record()
            self.end.synchronize()
            self.duration = self.start.elapsed_time(self.end)
        else:
            self.duration = (time.perf_counter() - self.t0) * 1e3
        return False

    def get_duration(self) -> float:
        return self.duration


def bench_kineto(fn, kernel_names: Union[str, Tuple[str, ...]], num_tests: int = 30,
                 suppress_kineto_output: bool = False, trace_path: str = None, flush_l2: bool = False):
    """Mean device time (seconds) of the kernels whose name contains ``kernel_names``, via torch.profiler."""
    fn()  # autotune / warm-up
    flush = torch.empty(int(256e6) // 4, dtype=torch.int32, device="cuda") if flush_l2 else None
    schedule = torch.profiler.schedule(wait=0, warmup=1, active=1, repeat=1)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA], schedule=schedule) as prof:
        for _ in range(2):
            for _ in range(num_tests):
                if flush is not None:
                    flush.zero_()
                fn()
            prof.step()
    is_tupled = isinstance(kernel_names, tuple)
    names = kernel_names if is_tupled else (kernel_names,)
    events = prof.key_averages()
    out = []
    for name in names:
        total_us = sum(e.device_time_total for e in events if name in e.key)
        # number of fn() calls the profiler actually recorded = launches of the kernel every call launches (the one with
        # the highest count); dividing by num_tests instead over-counts whenever the warm-up batch is traced as well
        calls = max((e.count for e in events if name in e.key), default=0)
        assert calls > 0, f"no profiled kernel matches '{name}'"
        out.append(total_us / calls / 1e6)
    if trace_path is not None:
        prof.export_chrome_trace(trace_path)
    return tuple(out) if is_tupled else out[0]


def GPU_bench(func, iters=100, warmup=30, kernel_name=None) -> float:
    """ms per call.  With ``kernel_name``: profiler time of the matching kernels with an L2 flush per
    iteration (the reference's convention for Voltrix, bench/bm_voltrix.py:36); otherwise CUDA events around
    ``iters`` back-to-back calls (its convention for cuSPARSE, bench/bm_sparse.py:27-42)."""
    if kernel_name is None:
        for _ in range(warmup):
            func()
        with DurationTimer() as t:
            for _ in range(iters):
                func()
        return t.get_duration() / iters
    return bench_kineto(func, kernel_name, num_tests=iters, suppress_kineto_output=True, flush_l2=True) * 1e3


def CPU_bench(func, iters=100, warmup=30) -> float:
    for _ in range(warmup):
        func()
    start = time.time()
    for _ in range(iters):
        func()
    return (time.time() - start) * 1000 / iters
