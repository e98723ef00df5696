#!/usr/bin/env python
"""bench.py -- the headline benchmark: SpMM GFLOP/s (2*nnz*N) on the Reddit-shaped graph, N=128, fp16 in /
fp32 accumulate (BASELINE.json configs[1]), one B200 or row-sharded over N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload reddit|products|rmat25|c1]

A step is one SpMM  C = A @ B  over the whole graph (all ranks together).  `value` is whole-job GFLOP/s with
everything resident in HBM; `e2e` is the same metric through the public API (voltrix.spmm) with B coming from
pinned host memory and C read back to the host every step.  L2 is flushed between timed iterations (B, 59.6 MB,
would otherwise stay L2-resident from one iteration to the next).

`--impl reference`: the CPU SpMM of BASELINE.md (the reference has no CPU SpMM of its own; this is the oracle's
OpenMP C port, kind "port") timed on the box's host cores on a bounded row sample of the same graph.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "voltrix-spmm_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "SpMM GFLOP/s (2*nnz*N), N=128, fp16 in / fp32 acc"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------- workloads
def make_workload(name: str, device, scale: float):
    from voltrix import graphs
    if name == "reddit":
        indptr, indices = graphs.reddit_shaped(seed=0, device=device, scale=scale)
        return indptr, indices, 128, "reddit-shaped Chung-Lu (M=232965, ~114.6M nnz) N=128 fp16"
    if name == "products":
        indptr, indices = graphs.products_shaped(seed=0, device=device, scale=scale)
        return indptr, indices, 256, "products-shaped Chung-Lu (M=2449029, ~123.7M nnz) N=256 fp16"
    if name == "rmat25":
        sc = 25 if scale >= 1 else max(12, int(25 + np.log2(scale)))
        indptr, indices = graphs.rmat_csr(sc, 32, seed=0, device=device)
        return indptr, indices, 256, f"R-MAT scale {sc} (0.57,0.19,0.19,0.05) edge factor 32, N=256 fp16"
    if name == "c1":
        indptr, indices = graphs.uniform_csr(16384, 1_000_000, seed=0, device=device)
        return indptr, indices, 64, "uniform 16384^2, 1M nnz, N=64 fp16"
    if name in {n for n, _, _ in graphs.named_suite()}:      # C3 shapes (scripts/ only; the bench CLI does not list them)
        indptr, indices = graphs.suite_graph(name, seed=0, device=device)
        return indptr, indices, 128, f"{name}-shaped Chung-Lu (C3 suite), N=128 fp16"
    raise SystemExit(f"unknown workload {name}")


def alg_bytes(nnz: int, M: int, K: int, N: int, in_bytes: int) -> int:
    """CSR indices once + indptr + B once + fp32 C once."""
    return 4 * nnz + 4 * (M + 1) + K * N * in_bytes + M * N * 4


def measured_traffic(workload: str, scale: float):
    """DRAM bytes per launch of the dominant kernel on this workload from the committed ncu capture (profiles/traffic.json:
    one `ncu --set full` launch after an L2 flush; the entry names the kernel variant it was taken on), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            db = json.load(f)
        e = db.get(workload if scale == 1.0 else f"{workload}@{scale:g}")
        if e is not None:
            return int(e["dram_bytes_read"] + e["dram_bytes_write"]), f"{e['kernel']}: {e['source']}"
    except Exception:
        pass
    return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(indptr_h: np.ndarray, indices_h: np.ndarray, K: int, N: int, budget_s: float = 12.0):
    """Oracle C port (OpenMP, all host cores) on a bounded row sample of the same graph."""
    import oracle
    c = oracle.c()
    c.set_num_threads(os.cpu_count() or 1)
    M = indptr_h.size - 1
    rng = np.random.default_rng(0)
    B = rng.random((K, N), dtype=np.float32)
    probe_rows = max(256, M // 64)
    t0 = time.perf_counter()
    c.spmm_csr(indptr_h, indices_h, B, 0, probe_rows, assume_coalesced=True)
    t_probe = time.perf_counter() - t0
    nnz_probe = int(indptr_h[probe_rows])
    rate = nnz_probe / max(t_probe, 1e-6)                          # nnz/s incl. thread spin-up
    want_nnz = min(int(indptr_h[-1]), int(rate * budget_s))
    rows = int(np.searchsorted(indptr_h, want_nnz, side="right")) - 1
    rows = max(probe_rows, min(M, rows))
    out = np.empty((rows, N), np.float32)
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter()
        c.spmm_csr(indptr_h, indices_h, B, 0, rows, assume_coalesced=True, out=out)
        best = min(best, time.perf_counter() - t0)
    nnz_s = int(indptr_h[rows])
    return {"value": 2.0 * nnz_s * N / best / 1e9, "unit": "GFLOP/s", "cores": c.num_threads(), "kind": "port",
            "sample": f"rows [0,{rows}) of {M} ({nnz_s} nnz), fp32, oracle/voltrix_oracle.c vo_spmm_csr (OpenMP), "
                      f"best of 2, {best * 1e3:.1f} ms",
            "host_cpus": os.cpu_count()}


def reference_preprocess_baseline(indptr_h, indices_h, ours_ms, budget_nnz=3_000_000):
    """The reference's OWN host preprocessing (voltrix::preprocess, bmat_kernels.cuh:264-320, compiled unmodified into
    oracle/_ref; one host thread as in the reference) on a bounded row prefix, beside the GPU preprocessing of this run."""
    import oracle
    M = indptr_h.size - 1
    rows = max(16, min(M, int(np.searchsorted(indptr_h, budget_nnz, side="right")) - 1) // 16 * 16)
    ip = np.ascontiguousarray(indptr_h[: rows + 1])
    ix = np.ascontiguousarray(indices_h[: ip[-1]])
    ref = oracle.ref()
    t0 = time.perf_counter()
    ref.preprocess(ip, ix)
    dt = time.perf_counter() - t0
    nnz_s, nnz = int(ix.size), int(indices_h.size)
    return {"kind": "reference (oracle/_ref: unmodified voltrix::preprocess, 1 host thread; stage a2 only, a3/a4 not included)",
            "sample": f"rows [0,{rows}) ({nnz_s} of {nnz} nnz)", "sample_ms": dt * 1e3,
            "extrapolated_whole_graph_ms": dt * 1e3 * nnz / max(nnz_s, 1),
            "this_repo_gpu_csr_preprocess_ms_whole_graph": ours_ms}


def extra_cpu_baselines(indptr_h, indices_h, K, N, budget_nnz=4_000_000):
    """scipy (1 thread) and torch.sparse CPU (all threads) on a smaller sample, as BASELINE.md lists them."""
    import scipy.sparse as sp
    M = indptr_h.size - 1
    rows = max(1, min(M, int(np.searchsorted(indptr_h, budget_nnz, side="right")) - 1))
    ip, ix = indptr_h[: rows + 1], indices_h[: indptr_h[rows]]
    B = np.random.default_rng(0).random((K, N), dtype=np.float32)
    A = sp.csr_matrix((np.ones(ix.size, np.float32), ix, ip), shape=(rows, K))
    t0 = time.perf_counter(); A @ B; ts = time.perf_counter() - t0
    out = {"scipy_1thread_gflops": 2.0 * ix.size * N / ts / 1e9}
    try:
        At = torch.sparse_csr_tensor(torch.from_numpy(ip.astype(np.int64)), torch.from_numpy(ix.astype(np.int64)),
                                     torch.ones(ix.size), size=(rows, K))
        Bt = torch.from_numpy(B)
        At @ Bt
        t0 = time.perf_counter(); At @ Bt; tt = time.perf_counter() - t0
        out["torch_sparse_cpu_gflops"] = 2.0 * ix.size * N / tt / 1e9
        out["torch_threads"] = torch.get_num_threads()
    except Exception as e:  # pragma: no cover
        out["torch_sparse_cpu_error"] = str(e)[:100]
    out["sample"] = f"rows [0,{rows}) ({ix.size} nnz)"
    return out


# ------------------------------------------------------------------------------------------- GPU baselines
def gpu_baselines(indptr, indices, M, N, feat16, iters=3):
    """cuSPARSE (torch.sparse_csr @ dense, the reference's bench/bm_sparse.py protocol) and the reference's own
    kernel recompiled for sm_100a (oracle/_ref), on the same inputs.  Reported, not optimised."""
    out = {}
    nnz = indices.numel()
    flops = 2.0 * nnz * N

    def timeit(fn):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    try:
        csr = torch.sparse_csr_tensor(indptr, indices, torch.ones(nnz, device="cuda"), size=(M, M))
        f32 = feat16.float()
        ms = timeit(lambda: csr @ f32)
        out["cusparse_fp32_ms"] = ms
        out["cusparse_fp32_gflops"] = flops / ms / 1e6
        del csr
        try:
            csr16 = torch.sparse_csr_tensor(indptr, indices, torch.ones(nnz, device="cuda", dtype=torch.float16),
                                            size=(M, M))
            ms = timeit(lambda: csr16 @ feat16)
            out["cusparse_fp16_ms"] = ms
            out["cusparse_fp16_gflops"] = flops / ms / 1e6
            del csr16
        except Exception as e:
            out["cusparse_fp16_error"] = str(e)[:120]
    except Exception as e:
        out["cusparse_error"] = str(e)[:120]
    torch.cuda.empty_cache()
    return out


def ref_kernel_baseline(blk, packed, hind, M, nnz, N, feat16, iters=3):
    """The reference's Hopper-era kernel (spmm_kernels.cuh, models 0/1/2), compiled for sm_100a in oracle/_ref."""
    import oracle
    out = {}
    try:
        lib = oracle.ref().lib
        f32 = feat16.float().contiguous()
        o = torch.empty(M, N, device="cuda")
        best = None
        for model in (0, 1, 2):
            def fn():
                rc = lib.ref_spmm(blk.data_ptr(), packed.data_ptr(), hind.data_ptr(), M, nnz, N, f32.data_ptr(),
                                  o.data_ptr(), model, torch.cuda.current_stream().cuda_stream)
                assert rc == 0
            fn(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            out[f"ref_kernel_model{model}_ms"] = ms
            best = ms if best is None else min(best, ms)
        out["ref_kernel_sm100a_fp32_gflops"] = 2.0 * nnz * N / best / 1e6
    except Exception as e:
        out["ref_kernel_error"] = str(e)[:160]
    return out


# ------------------------------------------------------------------------------------------- arms
def workload_config(desc: str, M: int, nnz: int, N: int) -> dict:
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": desc, "M": M, "nnz": nnz, "N": N,
            "l2": "GPU arm: flushed (256 MB write) before every timed step; CPU arm: operands exceed the host caches"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    scale = args.scale if dev == "cuda" else min(args.scale, 0.05)
    indptr, indices, N, desc = make_workload(args.workload, dev, scale)
    M = indptr.numel() - 1
    indptr_h, indices_h = indptr.cpu().numpy(), indices.cpu().numpy()
    del indptr, indices
    import oracle
    c = oracle.c()
    c.set_num_threads(os.cpu_count() or 1)     # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    B = np.random.default_rng(0).random((M, N), dtype=np.float32)
    # bounded sample per step: ~ (budget / (steps + warmup)) seconds of CPU work
    per_step_s = max(1.0, 100.0 / (args.steps + args.warmup))
    t0 = time.perf_counter()
    probe_rows = max(256, M // 64)
    c.spmm_csr(indptr_h, indices_h, B, 0, probe_rows, assume_coalesced=True)
    rate = int(indptr_h[probe_rows]) / max(time.perf_counter() - t0, 1e-6)
    rows = int(np.searchsorted(indptr_h, min(int(indptr_h[-1]), int(rate * per_step_s)), side="right")) - 1
    rows = max(probe_rows, min(M, rows))
    out = np.empty((rows, N), np.float32)
    for _ in range(args.warmup):
        c.spmm_csr(indptr_h, indices_h, B, 0, rows, assume_coalesced=True, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.spmm_csr(indptr_h, indices_h, B, 0, rows, assume_coalesced=True, out=out)
    dt = (time.perf_counter() - t0) / args.steps
    nnz_s = int(indptr_h[rows])
    val = 2.0 * nnz_s * N / dt / 1e9
    sample = f"rows [0,{rows}) of {M} ({nnz_s} of {indices_h.size} nnz) per step, fp32"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(desc, M, int(indices_h.size), N),
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": c.num_threads(), "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(),
                             "note": "the reference ships no CPU SpMM; oracle/voltrix_oracle.c vo_spmm_csr (OpenMP)"},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def build_sharded(workload: str, scale: float, dev, world: int):
    """(ShardedSpMM, whole-graph indptr/indices or None, M, nnz, N, description).  The R-MAT of C5 is never materialised
    on one rank when the job is sharded: every rank replays the same seeded edge stream twice -- once for the row
    histogram the partition needs, once keeping only its own rows."""
    from voltrix import graphs
    from voltrix.distributed import ROW_COST, ShardedSpMM
    if workload == "rmat25" and world > 1:
        sc = 25 if scale >= 1 else max(12, int(25 + np.log2(scale)))
        hist = graphs.rmat_row_histogram(sc, 32, seed=0, device=dev)          # draws per row (before coalescing)
        sh = ShardedSpMM(None, None, 1 << sc, weights=hist + ROW_COST,
                         local_csr=lambda r0, r1: graphs.rmat_csr(sc, 32, seed=0, device=dev, row_range=(r0, r1)))
        del hist
        nnz_t = torch.tensor([sh.local_nnz], device=dev, dtype=torch.int64)
        import torch.distributed as dist
        dist.all_reduce(nnz_t)
        return sh, None, None, 1 << sc, int(nnz_t.item()), 256, \
            f"R-MAT scale {sc} (0.57,0.19,0.19,0.05) edge factor 32, N=256 fp16"
    indptr, indices, N, desc = make_workload(workload, dev, scale)   # same seeded graph on every rank
    M = indptr.numel() - 1
    return ShardedSpMM(indptr, indices, M), indptr, indices, M, indices.numel(), N, desc


def parity_check(indptr_h, indices_h, row0, feat, out_rows, budget_nnz: int, ncols: int):
    """The timed GPU result against the CPU oracle (vo_spmm_csr, fp32 sums of the same fp16-rounded operand) on a bounded
    row prefix of rank 0's shard: the comparison of the reference's tests/test_spmm.py:75-96, inside the bench run.
    ``indptr_h`` / ``indices_h``: the CSR rows the GPU result covers (local numbering); ``out_rows``: GPU C, same rows."""
    import oracle
    from voltrix.utils import calc_diff, relative_error
    c = oracle.c()
    c.set_num_threads(os.cpu_count() or 1)
    M = indptr_h.size - 1
    rows = max(1, min(M, int(np.searchsorted(indptr_h, budget_nnz, side="right")) - 1))
    B = feat[:, :ncols].float().cpu().numpy()     # output columns are independent: a column slice is a valid check
    # double accumulator, rounded once: hub rows of the R-MAT sum millions of terms, where a SEQUENTIAL fp32 sum is
    # itself off by ~1e-3 (the rounding step of a 2^21-sized partial sum is 0.25) -- that error is not the kernel's
    want = torch.from_numpy(c.spmm_csr(indptr_h, indices_h, B, 0, rows, assume_coalesced=True, acc64=True))
    got = out_rows[:rows, :ncols].float().cpu()
    scale = max(float(want.abs().max()), 1e-9)
    err = float((got - want).abs().max()) / scale
    cd = float(calc_diff(got, want))
    rel = float(relative_error(got, want))
    return {"rows_checked": rows, "first_row": row0, "cols_checked": ncols, "nnz_checked": int(indptr_h[rows]),
            "max_scaled_err": err,
            "calc_diff": cd, "difference_rate_pct": f"{cd * 100:.2f}", "relative_error": rel,
            "ok": bool(err <= 1e-4 and rel <= 1e-2 and f"{cd * 100:.2f}" in ("0.00", "-0.00")),
            "oracle": "oracle/voltrix_oracle.c vo_spmm_csr_acc64 (fp64 accumulate, rounded once) on the same fp16-rounded "
                      "operand"}


def host_link_floor(world: int, h2d_bytes: int, d2h_bytes: int):
    """What the host<->device links of this pool's boxes allow for the end-to-end leg (measured, profiles/host_links.json):
    the slower direction of a step at the aggregate bandwidth `world` concurrently copying GPUs reach -- with that direction
    alone on the links (a floor) and with both directions busy (what a fully overlapped pipeline sees most of the time)."""
    try:
        with open(os.path.join(ROOT, "profiles", "host_links.json")) as f:
            db = json.load(f)
        k = str(world)
        if k not in db["d2h"]:
            return None
        gib = float(2**30)
        alone = max(d2h_bytes / (db["d2h"][k] * gib), h2d_bytes / (db["h2d"][k] * gib)) * 1e3      # each direction by itself
        duplex = max(d2h_bytes, h2d_bytes) / (db["both"][k] * gib) * 1e3                            # both directions busy
        return {"ms_per_step": alone, "ms_per_step_both_directions_busy": duplex,
                "aggregate_gib_s": {"d2h": db["d2h"][k], "h2d": db["h2d"][k], "both_per_direction": db["both"][k]},
                "source": db["source"]}
    except Exception:
        return None


def committed_floors(workload: str, scale: float, world: int, ms: float):
    """Measured floors of the tensor-core kernel (timing-only builds) from the committed report, if it covers this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "floors.json")) as f:
            db = json.load(f)
        e = db.get(workload)
        if e is None or scale != 1.0 or world != 1:
            return None
        out = dict(e)
        out["frac_of_binding_floor"] = max(v for k, v in e.items() if k.endswith("_ms")) / ms
        return out
    except Exception:
        return None


def measure(workload: str, args, world: int, rank: int, dev, headline: bool, steps: int, warmup: int):
    """One workload, all ranks: generate, shard, preprocess, time `steps` SpMMs (device events, L2 flushed, max over
    ranks), check the timed result against the oracle, and (headline or small operands) the end-to-end leg."""
    import torch.distributed as dist
    import voltrix

    t_gen = time.perf_counter()
    sh, indptr, indices, M, nnz, N, desc = build_sharded(workload, args.scale, dev, world)
    torch.cuda.synchronize()
    log(f"[rank {rank}] {workload}: M={M} nnz={nnz} N={N} generated + preprocessed in {time.perf_counter() - t_gen:.1f}s")
    blk, packed, hind = sh.state
    plan = packed._vx_plan
    # preprocessing alone, timed on a second pass when the whole graph is at hand (excluded from the step, like
    # bench/bm_voltrix.py:17 vs :36)
    t_pre = None
    if indptr is not None and headline:
        from voltrix.distributed import shard_csr
        lp, li = shard_csr(indptr, indices, sh.r0, sh.r1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        voltrix.csr_preprocess(lp, li, sh.local_rows, num_cols=M)
        torch.cuda.synchronize(); t_pre = time.perf_counter() - t0
        del lp, li
    log(f"[rank {rank}] rows [{sh.r0},{sh.r1}) nnz={sh.local_nnz} TCB={plan.total_blocks} items={plan.num_items} "
        f"sparse_rows={plan.num_sparse_rows} fixups={plan.num_fixups}"
        + (f" preprocess={t_pre * 1e3:.1f} ms" if t_pre else ""))

    # --- dense operand: created on rank 0, broadcast over NCCL (the path's one exchange step) ---
    g = torch.Generator(device=dev).manual_seed(0)
    feat = torch.rand(M, N, device=dev, generator=g).half() if rank == 0 else torch.empty(M, N, device=dev,
                                                                                         dtype=torch.float16)
    t_bcast = 0.0
    if world > 1:
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        sh.broadcast_features(feat, src=0)
        torch.cuda.synchronize(); t_bcast = time.perf_counter() - t0
    out = torch.empty(sh.local_rows, N, device=dev)

    def step():
        return voltrix.spmm(blk, packed, hind, sh.local_rows, sh.local_nnz, feat, out=out)

    step(); torch.cuda.synchronize()     # autotune + JIT load
    from voltrix.jit_kernels.spmm import feature_hash
    fh = feature_hash(packed)
    tuned = {str(k): v for k, v in voltrix.jit_tuner.tuned_keys.items()
             if k[0] == "spmm_kernel" and f"'N': {N}," in k[1] and f"'{fh}'" in k[1]}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        flush.zero_(); step()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    with clocks:
        barrier(); t_wall = time.perf_counter()
        for i in range(steps):
            flush.zero_()
            starts[i].record(); step(); ends[i].record()
        barrier(); t_wall = time.perf_counter() - t_wall
        if headline:
            # keep the sampler alive over ~1.5 s of back-to-back steps so nvidia-smi sees the kernel under load
            t_end = time.perf_counter() + 1.5
            while time.perf_counter() < t_end:
                for _ in range(20):
                    step()
                torch.cuda.synchronize()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    ms = float(np.mean(step_ms))
    log(f"[rank {rank}] {workload}: timed loop done: {ms:.3f} ms/step")
    per_rank = torch.zeros(world, device=dev, dtype=torch.float64)
    per_rank[rank] = ms
    if world > 1:
        dist.all_reduce(per_rank)
    per_rank_ms = [float(v) for v in per_rank.tolist()]
    ms_max = max(per_rank_ms)

    # --- parity of the TIMED result (rank 0's rows) against the CPU oracle, bounded ---
    parity = None
    if rank == 0:
        try:
            flush.zero_(); step(); torch.cuda.synchronize()
            if indptr is not None:
                lo = int(indptr[sh.r0])
                budget = 150_000_000      # the whole shard on C2 / C4 (~0.5-1 s of host time)
                rows_cap = int(torch.searchsorted(indptr[sh.r0:sh.r1 + 1] - lo, budget).item())
                rows_cap = max(1, min(sh.local_rows, rows_cap + 1))
                ip_h = (indptr[sh.r0:sh.r0 + rows_cap + 1] - lo).cpu().numpy().astype(np.int32)
                ix_h = indices[lo:lo + int(ip_h[-1])].cpu().numpy()
            else:
                ip_h, ix_h = plan.csr_indptr.cpu().numpy(), None
                rows_cap = max(1, min(sh.local_rows, int(np.searchsorted(ip_h, 50_000_000))))
                ip_h = np.ascontiguousarray(ip_h[: rows_cap + 1])
                ix_h = plan.csr_indices[: int(ip_h[-1])].cpu().numpy()
            parity = parity_check(ip_h, ix_h, sh.r0, feat, out, budget_nnz=int(ip_h[-1]),
                                  ncols=N if M * N * 4 <= (4 << 30) else 32)
            log(f"[rank 0] {workload}: parity {parity}")
        except Exception as ex:
            parity = {"error": str(ex)[:200], "ok": False}

    # --- e2e: public API with host buffers; H2D of B and D2H of C inside the timed region, every step ---
    # voltrix.HostStreamedSpMM runs copy-in, kernel and copy-out of consecutive steps on three streams (double-buffered).
    # Multi-GPU: every rank uploads only ITS 1/world row slice of B from pinned host memory and the ranks all-gather the
    # slices over NVLink (NCCL) on the copy-in stream -- B crosses the host links once per step, not once per rank -- and
    # every rank reads back its own rows of C.
    # Rank-INVARIANT decision (every rank takes the same branch: the legs below contain barriers).
    e2e_bytes = world * M * N * 2 + 2 * M * N * 4     # every rank pins B; C shards are pinned twice (double buffer)
    e2e = {"value": None, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    flops = 2.0 * nnz * N
    if e2e_bytes > (8 << 30) or not (headline or args.e2e_extras):
        e2e["skipped"] = (f"{e2e_bytes / 2**30:.0f} GiB of pinned host memory over {world} rank(s)" if e2e_bytes > (8 << 30)
                          else "extra workload: device-timed only")
    else:
        host_view = feat.cpu().pin_memory()     # the caller's B in pinned host memory (each rank reads only its slice)
        out_host = [torch.empty(sh.local_rows, N, dtype=torch.float32).pin_memory() for _ in range(2)]
        n_e2e = max(3, min(steps, 10))

        def time_e2e(run_steps):
            run_steps(2)
            barrier(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            run_steps(n_e2e)
            e.record(); barrier()
            t2 = torch.tensor([s.elapsed_time(e) / n_e2e], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            return float(t2.item())

        e2e_serial_ms = None
        if world == 1:
            feat_dev = torch.empty_like(feat)

            def serial_steps(n):
                for _ in range(n):
                    feat_dev.copy_(host_view, non_blocking=True)
                    o = voltrix.spmm(blk, packed, hind, sh.local_rows, sh.local_nnz, feat_dev, out=out)
                    out_host[0].copy_(o, non_blocking=True)

            e2e_serial_ms = time_e2e(serial_steps)
            del feat_dev
        pipe = voltrix.HostStreamedSpMM(blk, packed, hind, sh.local_rows, sh.local_nnz, N, dtype=feat.dtype, input_rows=M)

        def streamed_steps(n):
            pipe.fork()                                         # its streams start after the `s` event on this stream
            for i in range(n):
                pipe.submit(host_view, out_host[i % 2])
            pipe.join()                                         # this stream (and the `e` event) waits for the last D2H

        e2e_ms = time_e2e(streamed_steps)
        ok = bool(torch.equal(out_host[0], out_host[1]) and torch.equal(out_host[0], out.cpu()))
        e2e = {"value": flops / e2e_ms / 1e6, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(M * N * 2), "d2h_bytes_per_step": int(M * N * 4),
               "result_matches_device_run": ok,
               "api": "voltrix.HostStreamedSpMM.submit(pinned B, pinned C) every step: H2D of B (1/world row slice per rank "
                      "+ NCCL all-gather over NVLink when world > 1), voltrix.spmm, D2H of this rank's rows of C; the three "
                      "legs of consecutive steps overlap on three streams"}
        e2e["host_link_floor"] = host_link_floor(world, int(M * N * 2), int(M * N * 4))
        if e2e_serial_ms is not None:
            e2e["serial_ms_per_step"] = e2e_serial_ms
            e2e["serial_value"] = flops / e2e_serial_ms / 1e6
        del pipe, out_host

    res = {"workload": workload, "desc": desc, "M": M, "nnz": nnz, "N": N, "ms": ms, "ms_max": ms_max,
           "per_rank_ms": per_rank_ms, "step_ms": step_ms, "tuned": tuned, "t_pre": t_pre, "t_bcast": t_bcast,
           "t_wall": t_wall, "clocks": clocks.summary(), "parity": parity, "e2e": e2e, "sh": sh, "plan": plan,
           "feat": feat, "state": (blk, packed, hind), "indptr": indptr, "indices": indices, "flops": flops}
    return res


def roofline_of(r: dict, world: int, scale: float):
    M, N, nnz, sh, plan, ms = r["M"], r["N"], r["nnz"], r["sh"], r["plan"], r["ms"]
    peak, peak_src = measured_peak()
    abytes = alg_bytes(nnz, M, M, N, 2)
    # the dominant kernel on rank 0: algorithmic bytes of rank 0's shard / its launch duration
    abytes_local = 4 * sh.local_nnz + 4 * (sh.local_rows + 1) + M * N * 2 + sh.local_rows * N * 4
    achieved = abytes_local / (ms * 1e-3) / 1e9
    gather_bytes = (plan.total_blocks * 8 * N * 2 + 48 * plan.total_blocks + sh.local_rows * N * 4)
    traffic, traffic_src = measured_traffic(r["workload"], scale) if world == 1 else (None, None)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "alg_bytes_per_launch": abytes_local, "alg_bytes_whole_job": abytes,
            "gather_bytes_per_launch": gather_bytes, "gather_gbs": gather_bytes / (ms * 1e-3) / 1e9,
            "kernel_floors": committed_floors(r["workload"], scale, world, ms),
            "note": (f"B ({M * N * 2 / 1e6:.1f} MB fp16) is L2-resident: the kernel is bound by the L2->SM gather "
                     "stream (gather_bytes), not by compulsory HBM bytes -- see DESIGN.md section 4.5")
            if M * N * 2 < 100e6 else
                    (f"B ({M * N * 2 / 1e9:.2f} GB fp16) does not fit the 126 MB L2: the row gather "
                     "(gather_bytes, minus L2 hits on hub rows) is what HBM actually serves -- DESIGN.md 4.5")}


def run_product_arm(args):
    import faulthandler
    import torch.distributed as dist
    # hang diagnosis: dump every thread's Python stack to stderr if the bench is still running after this long
    faulthandler.dump_traceback_later(int(os.environ.get("VX_BENCH_STACK_DUMP_S", "1500")), repeat=False, exit=False)
    t_start = time.perf_counter()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (product arm) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a rank that stops making progress fails the collective after VX_BENCH_NCCL_TIMEOUT_S instead of 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(
            seconds=int(os.environ.get("VX_BENCH_NCCL_TIMEOUT_S", "600"))))

    r = measure(args.workload, args, world, rank, dev, headline=True, steps=args.steps, warmup=args.warmup)
    M, N, nnz, sh, plan, ms_max, ms = r["M"], r["N"], r["nnz"], r["sh"], r["plan"], r["ms_max"], r["ms"]
    flops = r["flops"]
    launches_per_step = 1 + (1 if plan.num_sparse_rows else 0) + (1 if plan.num_fixups else 0)   # kernels (+ one 4-byte memset)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC if N == 128 else METRIC.replace("N=128", f"N={N}"),
            "value": flops / ms_max / 1e6, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": workload_config(r["desc"], M, nnz, N),
            "details": {"sharding": f"{world} cost-balanced window-aligned row ranges" if world > 1 else "none",
                        "tuned": r["tuned"], "wall_ms_per_step_incl_flush": r["t_wall"] * 1e3 / args.steps,
                        "per_rank_ms": r["per_rank_ms"], "scheduler": "atomic-ticket claiming of LPT-sorted units, "
                        "feature-tile-major"},
            "roofline": roofline_of(r, world, args.scale),
            "parity": r["parity"],
            "e2e": r["e2e"],
            "gpu_launches": launches_per_step * args.steps,
            "clocks": r["clocks"],
            "preprocess_ms": r["t_pre"] * 1e3 if r["t_pre"] else None, "broadcast_ms": r["t_bcast"] * 1e3,
            "step_ms_min_med_max": [float(np.min(r["step_ms"])), float(np.median(r["step_ms"])), float(np.max(r["step_ms"]))],
        }
        if world == 1 and not args.no_baselines:
            try:
                blk, packed, hind = r["state"]
                line["baselines"] = gpu_baselines(r["indptr"], r["indices"], M, N, r["feat"])
                line["baselines"].update(ref_kernel_baseline(blk, packed, hind, M, nnz, N, r["feat"]))
            except Exception as ex:
                line["baselines"] = {"error": str(ex)[:200]}
        if world == 1:
            try:
                ip_h, ix_h = r["indptr"].cpu().numpy(), r["indices"].cpu().numpy()
                line["cpu_baseline"] = cpu_baseline(ip_h, ix_h, M, N)
                if not args.no_baselines:
                    line["cpu_baseline"]["others"] = extra_cpu_baselines(ip_h, ix_h, M, N)
                    try:
                        line["cpu_baseline"]["reference_preprocess"] = reference_preprocess_baseline(
                            ip_h, ix_h, r["t_pre"] * 1e3 if r["t_pre"] else None)
                    except Exception as ex:
                        line["cpu_baseline"]["reference_preprocess"] = {"error": str(ex)[:160]}
                del ip_h, ix_h
            except Exception as ex:
                line["cpu_baseline"] = {"error": str(ex)[:200]}
    del r
    torch.cuda.empty_cache()

    # --- the other north_star configurations, device-timed at this N (C4 products-shaped N=256; C5 R-MAT 2^25 rows) ---
    extras = [w for w in args.extra.split(",") if w and w != args.workload] if args.scale == 1.0 else []
    # An extra must never cost the headline: if one of them wedges (a rank lost inside a collective), every rank leaves
    # at the deadline and rank 0 prints the line with what it has.
    emitted = threading.Lock()

    def bail_out():
        if rank == 0 and emitted.acquire(blocking=False):
            line.setdefault("extra", {})["watchdog"] = f"extras cut off after {args.extra_budget_s + 240:.0f} s"
            emit(line)
        os._exit(0)

    watchdog = threading.Timer(args.extra_budget_s + 240, bail_out)
    watchdog.daemon = True
    if extras:
        watchdog.start()
    for w in extras:
        spent = torch.tensor([time.perf_counter() - t_start], device=dev)
        if world > 1:
            dist.all_reduce(spent, op=dist.ReduceOp.MAX)          # rank-invariant decision
        if float(spent.item()) > args.extra_budget_s:
            if rank == 0:
                line.setdefault("extra", {})[w] = {"skipped": f"time budget ({args.extra_budget_s:.0f} s) spent"}
            continue
        try:
            x = measure(w, args, world, rank, dev, headline=False, steps=min(args.steps, 10), warmup=3)
            if rank == 0:
                rf = roofline_of(x, world, args.scale)
                line.setdefault("extra", {})[w] = {
                    "config": workload_config(x["desc"], x["M"], x["nnz"], x["N"]),
                    "value": x["flops"] / x["ms_max"] / 1e6, "unit": "GFLOP/s", "ms_per_step": x["ms_max"],
                    "per_rank_ms": x["per_rank_ms"], "tuned": x["tuned"], "parity": x["parity"],
                    "roofline": {k: rf[k] for k in ("achieved", "peak", "frac", "traffic", "traffic_source",
                                                    "alg_bytes_whole_job", "gather_gbs")},
                    "e2e": x["e2e"], "clocks": x["clocks"],
                    "shard_rows": [b - a for a, b in x["sh"].ranges]}
            del x
        except Exception as ex:   # an extra must never cost the headline line
            log(f"[rank {rank}] extra workload {w} failed: {ex!r}")
            if rank == 0:
                line.setdefault("extra", {})[w] = {"error": repr(ex)[:300]}
            if world > 1:
                break             # ranks may have diverged inside a collective: stop issuing more
        torch.cuda.empty_cache()
    faulthandler.cancel_dump_traceback_later()
    watchdog.cancel()
    if rank == 0 and emitted.acquire(blocking=False):
        emit(line)
    if world > 1:
        try:
            dist.barrier(); dist.destroy_process_group()
        except Exception:
            pass


def emit(line: dict):
    """The ONE JSON line goes to the process's real stdout; everything else that writes to fd 1 while the bench runs
    (NCCL prints its version banner there) has been redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="voltrix", choices=["voltrix", "reference"])
    ap.add_argument("--workload", default="reddit", choices=["reddit", "products", "rmat25", "c1"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debug)")
    ap.add_argument("--no-baselines", action="store_true")
    ap.add_argument("--extra", default=os.environ.get("VX_BENCH_EXTRA", "products,rmat25"),
                    help="comma-separated workloads measured after the headline one and reported under `extra`")
    ap.add_argument("--extra-budget-s", type=float, default=float(os.environ.get("VX_BENCH_EXTRA_BUDGET_S", "420")),
                    help="no further extra workload is started once the run is this old")
    ap.add_argument("--e2e-extras", action="store_true", help="also run the host-buffer leg on the extra workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()
