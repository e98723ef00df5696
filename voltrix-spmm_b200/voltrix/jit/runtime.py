"""Loaded JIT artefacts.

A ``Runtime`` is a cache directory ``kernel.<name>.<hash>/{kernel.cu,kernel.args,kernel.so}`` whose
``kernel.so`` exports ``launch`` (same layout and call protocol as the reference,
voltrix/jit/runtime.py:9-72).  Calling it marshals the arguments with ctypes, passes a trailing
``int&`` and returns its value: 0 on success, a VX_* error code otherwise.
"""
import ctypes
import os
from typing import Optional

import torch

from .template import map_ctype


class Runtime:
    FILES = ("kernel.cu", "kernel.args", "kernel.so")

    def __init__(self, path: str) -> None:
        self.path = path
        self.lib = None
        self.args = None
        assert self.is_path_valid(self.path), f"{path} is not a complete JIT artefact"

    @staticmethod
    def is_path_valid(path: str) -> bool:
        return os.path.isdir(path) and all(os.path.exists(os.path.join(path, f)) for f in Runtime.FILES)

    def _load(self) -> None:
        self.lib = ctypes.CDLL(os.path.join(self.path, "kernel.so"))
        with open(os.path.join(self.path, "kernel.args"), "r") as f:
            self.args = eval(f.read())  # noqa: S307 -- written by build(); a tuple list of (name, type)
        if self.args and not isinstance(self.args[0], tuple):
            self.args = (self.args,)  # single-argument kernels: "('x', int)" evals to one tuple

    def __call__(self, *args) -> int:
        if self.lib is None or self.args is None:
            self._load()
        assert len(args) == len(self.args), f"Expected {len(self.args)} arguments, got {len(args)}"
        cargs = []
        for arg, (name, dtype) in zip(args, self.args):
            if arg is None:
                pass  # optional tensor -> null pointer
            elif isinstance(arg, torch.Tensor):
                assert arg.dtype == dtype, f"Expected tensor dtype `{dtype}` for `{name}`, got `{arg.dtype}`"
            else:
                assert isinstance(arg, dtype), f"Expected built-in type `{dtype}` for `{name}`, got `{type(arg)}`"
            cargs.append(map_ctype(arg, dtype))
        return_code = ctypes.c_int(0)
        self.lib.launch(*cargs, ctypes.byref(return_code))
        return return_code.value


class RuntimeCache:
    def __init__(self) -> None:
        self.cache = {}

    def __getitem__(self, path: str) -> Optional[Runtime]:
        if path in self.cache:
            return self.cache[path]
        if Runtime.is_path_valid(path):
            self.cache[path] = Runtime(path)
            return self.cache[path]
        return None

    def __setitem__(self, path: str, runtime: Runtime) -> None:
        self.cache[path] = runtime
