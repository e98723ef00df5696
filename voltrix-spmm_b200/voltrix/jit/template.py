"""Code generation for JIT kernels: one ``extern "C" void launch(...)`` per kernel variant.

Interface follows the reference (voltrix/jit/template.py:12-135): ``generate(includes, arg_defs,
body)`` returns CUDA source whose ``launch`` takes the ``arg_defs`` in order (tensors and streams as
``void*``, ``int`` as 32-bit int, ``bool`` as bool) followed by ``int& __return_code``.  Differences:
fp16 / int64 / uint8 tensors are accepted, ``None`` is marshalled as a null pointer for optional
tensors, and the generated body is expected to *write* ``__return_code`` (the reference declares it
but never sets it).
"""
import ctypes
import os
from typing import Any, Dict, Iterable, Tuple

import torch

from ..project import DEBUG_FLAG, PROJECT_NAME_FULL

# (python-side name for kernel.args, C type in the signature, C type after the cast)
_TENSOR_TYPES = {
    torch.int32: ("torch.int", "int*"),
    torch.uint32: ("torch.uint32", "uint32_t*"),
    torch.int64: ("torch.int64", "int64_t*"),
    torch.uint8: ("torch.uint8", "uint8_t*"),
    torch.float32: ("torch.float", "float*"),
    torch.float16: ("torch.float16", "__half*"),
    torch.bfloat16: ("torch.bfloat16", "__nv_bfloat16*"),
    torch.float8_e4m3fn: ("torch.float8_e4m3fn", "__nv_fp8_e4m3*"),
}

# Name map for Python `eval` of kernel.args
typename_map: Dict[Any, str] = {
    bool: "bool",
    int: "int",
    float: "float",
    torch.cuda.Stream: "torch.cuda.Stream",
    **{t: v[0] for t, v in _TENSOR_TYPES.items()},
}

# ctypes used to marshal each argument
ctype_map: Dict[Any, Any] = {
    bool: ctypes.c_bool,
    int: ctypes.c_int,
    float: ctypes.c_float,
    torch.cuda.Stream: ctypes.c_void_p,
    **{t: ctypes.c_void_p for t in _TENSOR_TYPES},
}

# (type in the extern "C" signature, type the body sees)
genc_map: Dict[Any, Tuple[str, str]] = {
    bool: ("bool", "bool"),
    int: ("int", "int"),
    float: ("float", "float"),
    torch.cuda.Stream: ("void*", "cudaStream_t"),
    **{t: ("void*", v[1]) for t, v in _TENSOR_TYPES.items()},
}


def map_ctype(value: Any, declared: Any = None) -> Any:
    """Python value -> ctypes value.  ``None`` is a null pointer for a pointer-typed argument."""
    if value is None:
        assert declared is not None and ctype_map[declared] is ctypes.c_void_p, "None only for pointer arguments"
        return ctypes.c_void_p(None)
    if isinstance(value, torch.Tensor):
        return ctype_map[value.dtype](value.data_ptr())
    if isinstance(value, torch.cuda.Stream):
        return ctypes.c_void_p(value.cuda_stream)
    return ctype_map[type(value)](value)


def cpp_format(template: str, keys: Dict[str, Any]) -> str:
    """Substitute ``{key}`` markers; C++ braces are left alone (str.format would choke on them)."""
    out = template
    for key, value in keys.items():
        out = out.replace(f"{{{key}}}", f"{value}")
    return out


def generate(includes: Iterable[str], arg_defs: Iterable[Tuple], body: str) -> str:
    assert isinstance(includes, (list, tuple))
    arg_defs = tuple(arg_defs)
    system = sorted({"<cuda.h>", "<cuda_runtime.h>", "<cuda_fp16.h>", "<cuda_bf16.h>", "<cuda_fp8.h>", "<cstdint>",
                     *[i for i in includes if i.startswith("<")]})
    package = sorted({i for i in includes if i.startswith('"')})
    lines = [f"// {PROJECT_NAME_FULL} (B200) auto-generated JIT CUDA source file", ""]
    lines += [f"#include {i}" for i in system] + [""]
    lines += [f"#include {i}" for i in package] + [""]

    params = []
    casts = []
    for name, ty in arg_defs:
        sig_t, body_t = genc_map[ty]
        if sig_t != body_t:
            params.append(f"{sig_t} __raw_{name}")
            casts.append(f"    auto {name} = reinterpret_cast<{body_t}>(__raw_{name});")
        else:
            params.append(f"{sig_t} {name}")
    params.append("int& __return_code")
    lines.append(f'extern "C" void launch({", ".join(params)}) {{')
    lines.append("    // Cast raw types (if needed)")
    lines += casts
    lines += [("    " + ln) if ln else "" for ln in body.split("\n")]
    lines += ["}", ""]
    code = "\n".join(lines)
    if os.getenv(DEBUG_FLAG, None):
        print(f"Generated code:\n{code}")
    return code
