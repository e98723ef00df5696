"""Compile-and-autotune front end of the JIT kernels.

Same entry point and semantics as the reference (voltrix/jit_kernels/tuner.py:42-168):
``jit_tuner.compile_and_tune(name, keys, space, includes, arg_defs, template, args, kernel_tag)``
formats ``template`` once per point of ``space``, builds every candidate, drops the ones that fail to
compile or return a non-zero code, times the rest and caches the winner for ``(name, keys)``.

Differences (all from SURVEY.md section 7.4):
* candidates are built from a thread pool that only shells out to nvcc -- no fork of a process that
  already holds a CUDA context (Q11);
* candidates are timed with CUDA events on the current stream with a 256 MB L2 flush between
  iterations (the reference's kineto path needs exactly one profiler row named like the kernel);
* winners are persisted in ``<cache>/tuned.json`` (the reference forgets them at exit), keyed by
  name, keys and GPU name;
* callers put N and dtype in ``keys`` (Q5).
"""
import copy
import json
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, List, Optional, Tuple

import torch

from ..jit import Runtime, build, cpp_format, generate
from ..jit.compiler import get_default_user_dir, put
from ..project import DEBUG_FLAG, PRINT_AUTOTUNE_FLAG


def _build_one(name, arg_defs, code, tuned_keys):
    try:
        return build(name, arg_defs, code), tuned_keys, None
    except Exception as e:  # a candidate that does not compile is simply not a candidate
        return None, tuned_keys, e


class JITTuner:
    def __init__(self) -> None:
        self.tuned: Dict[Tuple[str, str], Runtime] = {}
        self.tuned_keys: Dict[Tuple[str, str], dict] = {}
        self._disk: Optional[dict] = None

    # ------------------------------------------------------------------ persistence
    @staticmethod
    def _db_path() -> str:
        return os.path.join(get_default_user_dir(), "tuned.json")

    def _load_db(self) -> dict:
        if self._disk is None:
            try:
                with open(self._db_path(), "r") as f:
                    self._disk = json.load(f)
            except (OSError, ValueError):
                self._disk = {}
        return self._disk

    def _store_db(self, key: str, tuned_keys: dict, time_ms: float) -> None:
        db = self._load_db()
        db[key] = {"tuned_keys": tuned_keys, "time_ms": time_ms}
        try:
            put(self._db_path(), json.dumps(db, indent=1, sort_keys=True))
        except OSError:
            pass

    @staticmethod
    def _device_tag() -> str:
        return torch.cuda.get_device_name() if torch.cuda.is_available() else "nogpu"

    # ------------------------------------------------------------------ building
    @staticmethod
    def candidates(keys: dict, space: tuple, includes: tuple, arg_defs: tuple, template: str) -> List[Tuple[str, dict]]:
        space = (dict(),) if len(space) == 0 else space
        out = []
        for tuned_keys in space:
            assert isinstance(tuned_keys, dict)
            full_keys = copy.deepcopy(keys)
            full_keys.update(tuned_keys)
            out.append((generate(includes, arg_defs, cpp_format(template, full_keys)), tuned_keys))
        return out

    def precompile(self, name: str, keys: dict, space: tuple, includes: tuple, arg_defs: tuple, template: str):
        """Build every candidate without running anything (used by __graft_entry__.build on a GPU-less box)."""
        cands = self.candidates(keys, space, includes, arg_defs, template)
        with ThreadPoolExecutor(max_workers=min(len(cands), os.cpu_count() or 1)) as pool:
            results = list(pool.map(lambda c: _build_one(name, arg_defs, c[0], c[1]), cands))
        for runtime, tuned_keys, err in results:
            if runtime is None:
                raise RuntimeError(f"JIT kernel {name} {tuned_keys} failed to build: {err}")
        return [r for r, _, _ in results]

    # ------------------------------------------------------------------ tuning
    @staticmethod
    def _time(runtime: Runtime, args: tuple, iters: int = 8) -> float:
        flush = torch.empty(int(256e6) // 4, dtype=torch.int32, device="cuda")
        start = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
        end = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
        for i in range(iters):
            flush.zero_()
            start[i].record()
            rc = runtime(*args)
            end[i].record()
            if rc != 0:
                return float("inf")
        torch.cuda.synchronize()
        times = sorted(s.elapsed_time(e) for s, e in zip(start, end))
        return times[len(times) // 2]

    def compile_and_tune(self, name: str, keys: Dict[str, Any], space: tuple, includes: tuple, arg_defs: tuple,
                         template: str, args: tuple, kernel_tag=None) -> Runtime:
        keys = {k: keys[k] for k in sorted(keys.keys())}
        signature = (name, f"{keys}")
        if signature in self.tuned:
            if os.getenv(DEBUG_FLAG, None):
                print(f"Using cached JIT kernel {name} with keys {keys}")
            return self.tuned[signature]
        assert args is not None
        cands = self.candidates(keys, space, includes, arg_defs, template)

        # a persisted winner short-circuits both the other builds and the timing runs
        db_key = f"{name}|{keys}|{self._device_tag()}"
        if len(cands) > 1 and db_key in self._load_db():
            want = self._load_db()[db_key]["tuned_keys"]
            for code, tuned_keys in cands:
                if tuned_keys == want:
                    runtime, _, err = _build_one(name, arg_defs, code, tuned_keys)
                    # the persisted winner gets the same validity run a fresh candidate gets: if it cannot run these
                    # arguments (a matrix without the CSR arrays under a shared hash_tag, ...) fall through and re-tune
                    if runtime is not None and runtime(*args) == 0:
                        self.tuned[signature] = runtime
                        self.tuned_keys[signature] = tuned_keys
                        return runtime

        with ThreadPoolExecutor(max_workers=min(len(cands), os.cpu_count() or 1)) as pool:
            results = list(pool.map(lambda c: _build_one(name, arg_defs, c[0], c[1]), cands))
        kernels = [(r, k) for r, k, _ in results if r is not None]
        if os.getenv(DEBUG_FLAG, None):
            for _, k, err in results:
                if err is not None:
                    print(f"JIT kernel {name} candidate {k} failed to build: {err}")

        best_runtime, best_time, best_keys = None, None, None
        for runtime, tuned_keys in kernels:
            if len(cands) > 1:
                return_code = runtime(*args)  # validity run: unsupported configs report a non-zero code
                if return_code != 0:
                    if os.getenv(DEBUG_FLAG, None):
                        print(f"Illegal JIT kernel {name} keys {keys} tuned {tuned_keys}: code {return_code}")
                    continue
                elapsed = self._time(runtime, args)
            else:
                elapsed = 0.0
            if best_time is None or elapsed < best_time:
                best_runtime, best_time, best_keys = runtime, elapsed, tuned_keys
            if os.getenv(DEBUG_FLAG, None):
                print(f"Tuned JIT kernel {name} keys {keys} tuned {tuned_keys}: {elapsed:.4f} ms")
        assert best_runtime is not None, f"Failed to tune JIT kernel {name} with keys {keys}"
        if os.getenv(DEBUG_FLAG, None) or os.getenv(PRINT_AUTOTUNE_FLAG, None):
            print(f"JIT kernel {name}[{len(kernels)}/{len(cands)}] keys {keys} -> {best_keys} ({best_time:.4f} ms)")
        if len(cands) > 1:
            self._store_db(db_key, best_keys, best_time)
        self.tuned[signature] = best_runtime
        self.tuned_keys[signature] = best_keys
        return best_runtime


jit_tuner = JITTuner()
