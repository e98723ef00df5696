"""JIT layer: source generation, nvcc build + on-disk cache, ctypes loading (reference: voltrix/jit/__init__.py:1-3
re-exports the same five names)."""
from .runtime import Runtime, RuntimeCache
from .template import cpp_format, generate
from .compiler import build, get_nvcc_compiler

__all__ = ["get_nvcc_compiler", "build", "cpp_format", "generate", "Runtime", "RuntimeCache"]
