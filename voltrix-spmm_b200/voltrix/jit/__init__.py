from .compiler import get_nvcc_compiler, build
from .template import cpp_format, generate
from .runtime import Runtime
