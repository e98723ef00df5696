"""Loaded JIT artefacts.

An artefact is a directory ``kernel.<name>.<hash>/`` holding ``kernel.cu`` (generated source), ``kernel.args`` (the
``arg_defs`` as a Python literal) and ``kernel.so`` (exports ``extern "C" void launch(..., int& __return_code)``) -- the
layout and call protocol of the reference (voltrix/jit/runtime.py:9-72).  ``Runtime(path)(*args)`` checks the arguments
against ``kernel.args``, marshals them through ctypes and returns the code the kernel wrote: 0, or a VX_* error.
"""
import ast
import ctypes
import os
from typing import Dict, Optional, Sequence, Tuple

import torch

from .template import map_ctype

_ARTEFACT_FILES = ("kernel.cu", "kernel.args", "kernel.so")
# names that may appear in kernel.args (written by compiler.build from template.typename_map)
_ARG_NAMESPACE = {"torch": torch, "int": int, "bool": bool, "float": float}


def _read_arg_defs(path: str) -> Tuple[Tuple[str, object], ...]:
    """kernel.args is ``(('name', type), ...)``; a single-argument kernel writes one bare pair."""
    with open(path, "r") as fh:
        text = fh.read()
    ast.parse(text, mode="eval")                      # a literal expression, nothing else
    defs = eval(text, {"__builtins__": {}}, dict(_ARG_NAMESPACE))  # noqa: S307
    if defs and not isinstance(defs[0], tuple):
        defs = (defs,)
    return tuple(defs)


class Runtime:
    def __init__(self, path: str) -> None:
        if not self.is_path_valid(path):
            raise AssertionError(f"{path} is not a complete JIT artefact")
        self.path = path
        self._lib = None
        self._launch = None
        self._arg_defs: Optional[Tuple[Tuple[str, object], ...]] = None

    @staticmethod
    def is_path_valid(path: str) -> bool:
        return os.path.isdir(path) and all(os.path.isfile(os.path.join(path, name)) for name in _ARTEFACT_FILES)

    # kept for callers that poke at the loaded state like the reference's tests do
    @property
    def lib(self):
        return self._lib

    @property
    def args(self):
        return self._arg_defs

    def _ensure_loaded(self) -> None:
        if self._launch is None:
            self._lib = ctypes.CDLL(os.path.join(self.path, "kernel.so"))
            self._launch = self._lib.launch
            self._arg_defs = _read_arg_defs(os.path.join(self.path, "kernel.args"))

    def _marshal(self, values: Sequence) -> list:
        if len(values) != len(self._arg_defs):
            raise AssertionError(f"Expected {len(self._arg_defs)} arguments, got {len(values)}")
        out = []
        for value, (name, declared) in zip(values, self._arg_defs):
            if isinstance(value, torch.Tensor):
                if value.dtype != declared:
                    raise AssertionError(f"Expected tensor dtype `{declared}` for `{name}`, got `{value.dtype}`")
            elif value is not None and not isinstance(value, declared):   # None = null pointer for an optional tensor
                raise AssertionError(f"Expected built-in type `{declared}` for `{name}`, got `{type(value)}`")
            out.append(map_ctype(value, declared))
        return out

    def __call__(self, *values) -> int:
        self._ensure_loaded()
        code = ctypes.c_int(0)
        self._launch(*self._marshal(values), ctypes.byref(code))
        return code.value

    # -- prepared launches (extension): callers that issue the same launch again and again with only a few pointers
    # changing marshal the argument list once and patch the changing slots, instead of type-checking and converting
    # two dozen arguments per call (tens of microseconds of interpreter time -- more than a small SpMM takes on the GPU)
    def prepare(self, values: Sequence) -> list:
        """Type-check and marshal ``values`` once; the result goes to ``launch_prepared`` (entries may be replaced by
        ``ctypes.c_void_p`` / ``ctypes.c_int`` objects in between)."""
        self._ensure_loaded()
        return self._marshal(values)

    def arg_index(self, name: str) -> int:
        self._ensure_loaded()
        for i, (n, _) in enumerate(self._arg_defs):
            if n == name:
                return i
        raise KeyError(name)

    def launch_prepared(self, cvalues: list) -> int:
        code = ctypes.c_int(0)
        self._launch(*cvalues, ctypes.byref(code))
        return code.value


class RuntimeCache:
    """path -> Runtime; a path that is not (yet) a complete artefact maps to None."""

    def __init__(self) -> None:
        self._by_path: Dict[str, Runtime] = {}

    def __getitem__(self, path: str) -> Optional[Runtime]:
        hit = self._by_path.get(path)
        if hit is None and Runtime.is_path_valid(path):
            hit = self._by_path[path] = Runtime(path)
        return hit

    def __setitem__(self, path: str, runtime: Runtime) -> None:
        self._by_path[path] = runtime
