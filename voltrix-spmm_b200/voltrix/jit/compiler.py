"""nvcc driver and on-disk kernel cache.

Same contract as the reference (voltrix/jit/compiler.py:25-189): ``build(name, arg_defs, code)``
hashes (name, header contents, code, compiler, flags) into
``<cache>/kernel.<name>.<md5-12>/{kernel.cu,kernel.args,kernel.so}``, compiles with nvcc on a miss,
installs the .so atomically and returns a ``Runtime``.  What changed for B200:

* ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` instead of ``compute_90a`` (:125);
* the hash covers EVERY header under the include dir, not only ``*.cuh``;
* the nvcc version gate compares integers, not strings (Q10);
* the stray ``-I/home/...`` / ``-DDISABLE_MX_MAIN`` flags (:131-132) are gone;
* default cache dir is in-tree (``voltrix-spmm_b200/jit_cache``) so artefacts built on the CPU box
  travel to the GPU box; ``VOLTRIX_CACHE_DIR`` overrides it as in the reference.
"""
import functools
import hashlib
import os
import re
import subprocess
import uuid
from typing import Tuple, cast

from ..project import (CACHE_DIR_FLAG, DEBUG_FLAG, EXTRA_NVCC_FLAGS_FLAG, JIT_PRINT_NVCC_COMMAND_FLAG, NVCC_COMPILER_FLAG,
                       PROJECT_NAME_ABBR_LOWER, PTXAS_VERBOSE_FLAG)
from .runtime import Runtime, RuntimeCache
from .template import typename_map

runtime_cache = RuntimeCache()

SM_ARCH_FLAG = "-gencode=arch=compute_100a,code=sm_100a"


def hash_to_hex(s: str) -> str:
    return hashlib.md5(s.encode("utf-8")).hexdigest()[0:12]


@functools.lru_cache(maxsize=None)
def get_package_root() -> str:
    return os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


@functools.lru_cache(maxsize=None)
def get_jit_include_dir() -> str:
    return os.path.join(get_package_root(), "csrc")


@functools.lru_cache(maxsize=None)
def get_repo_version() -> str:
    """md5 over every header the JIT kernels can include."""
    include_dir = os.path.join(get_jit_include_dir(), PROJECT_NAME_ABBR_LOWER)
    assert os.path.isdir(include_dir), f"Cannot find include directory {include_dir}"
    md5 = hashlib.md5()
    for dirpath, _, filenames in sorted(os.walk(include_dir, followlinks=True)):
        for filename in sorted(filenames):
            if filename.endswith((".cuh", ".h", ".hpp")):
                with open(os.path.join(dirpath, filename), "rb") as f:
                    md5.update(filename.encode())
                    md5.update(f.read())
    return md5.hexdigest()[0:12]


def _cuda_home() -> str:
    home = os.environ.get("CUDA_HOME") or os.environ.get("CUDA_PATH")
    if home:
        return home
    try:
        from torch.utils.cpp_extension import CUDA_HOME
        if CUDA_HOME:
            return CUDA_HOME
    except Exception:  # pragma: no cover
        pass
    return "/usr/local/cuda"


@functools.lru_cache(maxsize=None)
def get_nvcc_compiler() -> Tuple[str, str]:
    paths = []
    if os.getenv(NVCC_COMPILER_FLAG):
        paths.append(os.getenv(NVCC_COMPILER_FLAG))
    paths.append(f"{_cuda_home()}/bin/nvcc")
    least = (12, 8)  # first toolkit with sm_100a
    pattern = re.compile(r"release (\d+)\.(\d+)")
    for path in paths:
        if os.path.exists(path):
            out = subprocess.run([path, "--version"], capture_output=True, text=True).stdout
            match = pattern.search(out)
            assert match, f"Cannot get the version of NVCC compiler {path}"
            version = (int(match.group(1)), int(match.group(2)))
            assert version >= least, f"NVCC {path} version {version} is lower than {least}"
            return path, f"{version[0]}.{version[1]}"
    raise RuntimeError("Cannot find any available NVCC compiler")


@functools.lru_cache(maxsize=None)
def get_default_user_dir() -> str:
    if CACHE_DIR_FLAG in os.environ:
        path = os.getenv(CACHE_DIR_FLAG)
    else:
        path = os.path.join(get_package_root(), "jit_cache")
    os.makedirs(path, exist_ok=True)
    return path


def get_tmp_dir() -> str:
    return f"{get_default_user_dir()}/tmp"


def get_cache_dir() -> str:
    return f"{get_default_user_dir()}/cache"


def make_tmp_dir() -> str:
    tmp_dir = get_tmp_dir()
    os.makedirs(tmp_dir, exist_ok=True)
    return tmp_dir


def put(path: str, data, is_binary: bool = False) -> None:
    """Write then POSIX-atomic replace, so concurrent builders never see a torn file."""
    tmp_file_path = f"{make_tmp_dir()}/file.tmp.{uuid.uuid4()}.{hash_to_hex(path)}"
    with open(tmp_file_path, "wb" if is_binary else "w") as f:
        f.write(data)
    os.replace(tmp_file_path, path)


def nvcc_flags() -> list:
    flags = [
        "-std=c++17",
        "-shared",
        "-O3",
        "-lineinfo",
        "--expt-relaxed-constexpr",
        "--expt-extended-lambda",
        SM_ARCH_FLAG,
        "--diag-suppress=177,174,940",
    ]
    if PTXAS_VERBOSE_FLAG in os.environ:
        flags.append("--ptxas-options=-v")
    flags += os.environ.get(EXTRA_NVCC_FLAGS_FLAG, "").split()   # experiments; part of the cache key
    cxx_flags = ["-fPIC", "-O3", "-Wno-deprecated-declarations", "-Wno-abi", "-fno-gnu-unique"]
    return [*flags, f'--compiler-options={",".join(cxx_flags)}']


def build(name: str, arg_defs: tuple, code: str) -> Runtime:
    flags = nvcc_flags()
    signature = f"{name}$${get_repo_version()}$${code}$${get_nvcc_compiler()}$${flags}"
    name = f"kernel.{name}.{hash_to_hex(signature)}"
    path = f"{get_cache_dir()}/{name}"

    global runtime_cache
    if runtime_cache[path] is not None:
        if os.getenv(DEBUG_FLAG, None):
            print(f"Using cached JIT runtime {name} during build")
        return cast(Runtime, runtime_cache[path])

    os.makedirs(path, exist_ok=True)
    put(f"{path}/kernel.args", ", ".join(f"('{n}', {typename_map[t]})" for n, t in arg_defs))
    src_path = f"{path}/kernel.cu"
    put(src_path, code)

    so_path = f"{path}/kernel.so"
    tmp_so_path = f"{make_tmp_dir()}/nvcc.tmp.{uuid.uuid4()}.{hash_to_hex(so_path)}.so"
    command = [get_nvcc_compiler()[0], src_path, "-o", tmp_so_path, *flags, f"-I{get_jit_include_dir()}"]
    if os.getenv(DEBUG_FLAG, None) or os.getenv(JIT_PRINT_NVCC_COMMAND_FLAG, False):
        print(f"Compiling JIT runtime {name} with command {command}")
    proc = subprocess.run(command, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"Failed to compile {src_path}:\n{proc.stdout}\n{proc.stderr}")
    if PTXAS_VERBOSE_FLAG in os.environ:
        print(proc.stderr)
    os.replace(tmp_so_path, so_path)

    runtime_cache[path] = Runtime(path)
    return cast(Runtime, runtime_cache[path])
