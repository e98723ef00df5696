import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "voltrix-spmm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _load_golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def golden_module():
    return _load_golden_module()


@pytest.fixture(scope="session")
def golden_cases(golden_module):
    """name -> dict(indptr, indices, full(bool), npz)"""
    out = {}
    for name, (indptr, indices), full in golden_module.cases():
        out[name] = dict(indptr=indptr, indices=indices, full=full, npz=np.load(os.path.join(GOLDEN, name + ".npz")))
    return out


def small_case_names():
    return ["tiny_37_unsorted_dups", "m100_empty_window", "m64_dense", "m1000_sparse", "m257_tail1_empty_tail",
            "m48_all_empty"]


def all_case_names():
    return small_case_names() + ["ref_test_spmm_kernel_seed20_d0.01", "c1_uniform_16384"]
