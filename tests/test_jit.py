"""JIT plumbing (reference: tests/test_jit.py:31-64): codegen + nvcc + cache + ctypes marshalling.
Runs without a GPU: the generated `launch` only echoes its arguments into a host tensor."""
import os

import pytest
import torch

import voltrix
from voltrix import jit
from voltrix.jit import compiler as jit_compiler
from voltrix.jit_kernels.tuner import JITTuner

ARG_DEFS = (
    ("out", torch.int64),          # host tensor the kernel writes into
    ("a", torch.float32),
    ("b", torch.bfloat16),
    ("c", torch.float16),
    ("d", torch.uint32),
    ("opt", torch.int32),          # passed as None -> null pointer
    ("n", int),
    ("flag", bool),
    ("scale", float),
)
BODY = """
out[0] = (int64_t)a; out[1] = (int64_t)b; out[2] = (int64_t)c; out[3] = (int64_t)d;
out[4] = (int64_t)opt; out[5] = n; out[6] = flag ? 1 : 0; out[7] = (int64_t)(scale * 4);
__return_code = {code};
"""


@pytest.fixture()
def cache_dir(tmp_path, monkeypatch):
    monkeypatch.setenv(voltrix.CACHE_DIR_FLAG, str(tmp_path))
    jit_compiler.get_default_user_dir.cache_clear()
    yield tmp_path
    jit_compiler.get_default_user_dir.cache_clear()


def test_generate_signature():
    code = jit.generate(('"voltrix/common.cuh"', "<vector>"), ARG_DEFS, jit.cpp_format(BODY, {"code": 0}))
    assert 'extern "C" void launch(void* __raw_out, void* __raw_a' in code
    assert "int n, bool flag, float scale, int& __return_code)" in code
    assert "auto b = reinterpret_cast<__nv_bfloat16*>(__raw_b);" in code
    assert '#include "voltrix/common.cuh"' in code and "#include <vector>" in code


def test_cpp_format_leaves_cpp_braces():
    assert jit.cpp_format("if (x) { f<{N}>(); }", {"N": 7}) == "if (x) { f<7>(); }"


def test_build_cache_and_marshalling(cache_dir):
    code = jit.generate(tuple(), ARG_DEFS, jit.cpp_format(BODY, {"code": 0}))
    rt = jit.build("echo", ARG_DEFS, code)
    assert isinstance(rt, jit.Runtime)
    files = sorted(os.listdir(rt.path))
    assert files == ["kernel.args", "kernel.cu", "kernel.so"]
    assert os.path.basename(rt.path).startswith("kernel.echo.") and str(cache_dir) in rt.path

    out = torch.zeros(8, dtype=torch.int64)
    a, b = torch.zeros(4), torch.zeros(4, dtype=torch.bfloat16)
    c, d = torch.zeros(4, dtype=torch.float16), torch.zeros(4, dtype=torch.uint32)
    assert rt(out, a, b, c, d, None, 42, True, 2.5) == 0
    assert out.tolist() == [a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), 0, 42, 1, 10]

    # second build of the same code is a cache hit (same object, no recompile)
    assert jit.build("echo", ARG_DEFS, code) is rt
    # wrong dtype / arity are rejected before the call
    with pytest.raises(AssertionError):
        rt(out, b, b, c, d, None, 42, True, 2.5)
    with pytest.raises(AssertionError):
        rt(out, a)


def test_return_code_reaches_python_and_tuner_skips_failing_candidates(cache_dir):
    tuner = JITTuner()
    out = torch.zeros(8, dtype=torch.int64)
    a, b = torch.zeros(4), torch.zeros(4, dtype=torch.bfloat16)
    c, d = torch.zeros(4, dtype=torch.float16), torch.zeros(4, dtype=torch.uint32)
    args = (out, a, b, c, d, None, 1, False, 1.0)
    # single-point space: no timing run, return code comes back from the call
    rt = tuner.compile_and_tune("echo_rc", {"k": 1}, ({"code": 3},), tuple(), ARG_DEFS, BODY, args)
    assert rt(*args) == 3
    # same (name, keys) -> in-process cache
    assert tuner.compile_and_tune("echo_rc", {"k": 1}, ({"code": 3},), tuple(), ARG_DEFS, BODY, args) is rt
    # a candidate that does not compile is dropped, the other one wins
    bad_good = ({"code": "this is not C++"}, {"code": 0})
    cands = tuner.candidates({"k": 2}, bad_good, tuple(), ARG_DEFS, BODY)
    assert len(cands) == 2 and "this is not C++" in cands[0][0]


def test_nvcc_compiler_version_is_numeric():
    path, version = jit.get_nvcc_compiler()
    assert os.path.exists(path)
    major, minor = (int(x) for x in version.split("."))
    assert (major, minor) >= (12, 8)


def test_repo_version_hashes_all_headers():
    v = jit_compiler.get_repo_version()
    assert len(v) == 12 and v == jit_compiler.get_repo_version()


def test_public_surface_matches_reference():
    # reference voltrix/__init__.py:1-3, jit/__init__.py:1-3, jit_kernels/__init__.py:1-4, spmm/__init__.py:1-5
    for name in ("BLK_H", "BLK_W", "csr_preprocess", "spmm", "preprocess_kernel", "hmat_gen_kernel",
                 "hmat_packed_swizzle_kernel", "spmm_kernel", "jit", "jit_kernels", "project",
                 "DEBUG_FLAG", "NVCC_COMPILER_FLAG", "CACHE_DIR_FLAG", "PTXAS_VERBOSE_FLAG",
                 "JIT_PRINT_NVCC_COMMAND_FLAG", "PRINT_AUTOTUNE_FLAG", "PROJECT_NAME_FULL", "PROJECT_NAME_ABBR",
                 "PROJECT_NAME_FULL_LOWER", "PROJECT_NAME_ABBR_LOWER"):
        assert hasattr(voltrix, name), name
    assert callable(voltrix.spmm) and (voltrix.BLK_H, voltrix.BLK_W) == (16, 8)
    for name in ("get_nvcc_compiler", "build", "cpp_format", "generate", "Runtime"):
        assert hasattr(voltrix.jit, name)
    from voltrix.utils import calc_diff, GPU_bench, relative_error, bench_kineto, DurationTimer  # noqa: F401


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        voltrix.csr_preprocess(torch.zeros(17, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), 16)


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "voltrix-spmm_b200")
    for dirpath, _, files in os.walk(root):
        if "jit_cache" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "voltrix_oracle" not in text, f


def test_environment_flag_spellings_are_the_references():
    # reference voltrix/project/const.py:2-14: the values, not just the names, are the contract (shells export them)
    assert (voltrix.PROJECT_NAME_FULL, voltrix.PROJECT_NAME_ABBR) == ("Voltrix-SpMM", "Voltrix")
    assert (voltrix.PROJECT_NAME_FULL_LOWER, voltrix.PROJECT_NAME_ABBR_LOWER) == ("voltrix-spmm", "voltrix")
    assert voltrix.DEBUG_FLAG == "VOLTRIX_JIT_DEBUG" and voltrix.NVCC_COMPILER_FLAG == "VOLTRIX_NVCC_COMPILER"
    assert voltrix.CACHE_DIR_FLAG == "VOLTRIX_CACHE_DIR" and voltrix.PTXAS_VERBOSE_FLAG == "VOLTRIX_PTXAS_VERBOSE"
    assert voltrix.JIT_PRINT_NVCC_COMMAND_FLAG == "VOLTRIX_JIT_PRINT_NVCC_COMMAND"
    assert voltrix.PRINT_AUTOTUNE_FLAG == "VOLTRIX_PRINT_AUTO_TUNE"
    assert voltrix.FP32_MODE_FLAG == "VOLTRIX_FP32_MODE" and voltrix.EXTRA_NVCC_FLAGS_FLAG == "VOLTRIX_EXTRA_NVCC_FLAGS"


def test_default_routing_rule_depends_on_size_only():
    """csr_preprocess's default (sparse_ratio, small_blocks): the small-window rule is on from SMALL_BLOCKS_MIN_TCB TC blocks
    up (it costs a second launch on small matrices); a plan without CSR arrays has no rule to deviate from."""
    from voltrix.spmm.spmm import SMALL_BLOCKS_MIN_TCB, SpmmPlan, default_routing
    assert default_routing(0) == (0.5, 0) and default_routing(SMALL_BLOCKS_MIN_TCB - 1) == (0.5, 0)
    assert default_routing(SMALL_BLOCKS_MIN_TCB) == (0.5, 8) and default_routing(10 ** 9) == (0.5, 8)
    assert voltrix.ROUTING_CANDIDATES[0] == (0.5, 0) and (0.5, 8) in voltrix.ROUTING_CANDIDATES
    plan = SpmmPlan()
    plan.total_blocks, plan.sparse_ratio, plan.small_blocks = 1000, 0.0, 0
    assert plan.route_is_default                        # no CSR arrays
    plan.csr_indptr = torch.zeros(2, dtype=torch.int32)
    assert not plan.route_is_default
    plan.sparse_ratio = 0.5
    assert plan.route_is_default
    plan.total_blocks = SMALL_BLOCKS_MIN_TCB
    assert not plan.route_is_default
    plan.small_blocks = 8
    assert plan.route_is_default
