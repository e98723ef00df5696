"""The C-ABI library loads and exports every symbol include/voltrix_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "voltrix_b200.h")
LIB = os.path.join(ROOT, "voltrix-spmm_b200", "csrc", "libvoltrix_b200.so")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build_capi()
    return ctypes.CDLL(LIB)


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("vx_preprocess", "vx_hmat_gen", "vx_hmat_packed_swizzle", "vx_csr_window_sort",
                 "vx_csr_tiles_scatter", "vx_schedule_build", "vx_schedule_sort", "vx_spmm"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in voltrix_b200.h but not exported"


def test_host_only_queries(lib):
    lib.vx_abi_version.restype = ctypes.c_int
    assert lib.vx_abi_version() == 5   # 5: vx_plan_t.ticket (dynamic unit claiming), value_tiles / csr_values + vx_value_tiles; 4: vx_spmm_csr_weighted; 2: vx_plan_t gained split_ws (fp32 tensor-core path); 3: fused epilogue fields
    lib.vx_preprocess_workspace_bytes.restype = ctypes.c_size_t
    lib.vx_preprocess_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int32]
    small = lib.vx_preprocess_workspace_bytes(1000, 100)
    big = lib.vx_preprocess_workspace_bytes(1_000_000, 16384)
    assert 0 < small < big and big >= 1_000_000 * 20
    lib.vx_schedule_max_items.restype = ctypes.c_int64
    lib.vx_schedule_max_items.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32]
    assert lib.vx_schedule_max_items(1600, 10_000, 64) >= 100 + 10_000 // 64


def test_struct_sizes_match_header():
    class Item(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int32) for n in ("window", "blk_begin", "blk_count", "slot")]
    assert ctypes.sizeof(Item) == 16
