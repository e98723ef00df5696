"""`pip install -e voltrix-spmm_b200` (reference: setup.py:6-11, `pip install -e .`): the editable install resolves to the
in-tree package, so the JIT finds its CUDA headers and the in-tree kernel cache."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "voltrix-spmm_b200")


def test_editable_install_imports_the_in_tree_package(tmp_path):
    target = tmp_path / "site"
    p = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "-q",
                        "-e", PKG, "--target", str(target)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    code = (
        "import site, sys, os\n"
        f"site.addsitedir({str(target)!r})\n"
        "import voltrix\n"
        "from voltrix.jit.compiler import get_jit_include_dir\n"
        "assert os.path.isfile(os.path.join(get_jit_include_dir(), 'voltrix', 'spmm_kernels.cuh'))\n"
        "for name in ('BLK_H', 'BLK_W', 'csr_preprocess', 'spmm', 'spmm_kernel', 'preprocess_kernel', 'hmat_gen_kernel',\n"
        "             'hmat_packed_swizzle_kernel', 'jit', 'jit_kernels', 'project', 'utils'):\n"
        "    assert hasattr(voltrix, name), name\n"
        "print(os.path.dirname(os.path.abspath(voltrix.__file__)))\n")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    q = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert q.returncode == 0, q.stderr[-2000:]
    assert q.stdout.strip().splitlines()[-1] == os.path.join(PKG, "voltrix")
