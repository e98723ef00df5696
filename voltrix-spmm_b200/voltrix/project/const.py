"""Project names and the environment variables this package reads.

The names and spellings the reference exports (voltrix/project/const.py:2-14) are kept, so that scripts and shells configured
for it keep working; the variables below the divider are new here.  Every variable is `VOLTRIX_` + a suffix.
"""
PROJECT_NAME_FULL, PROJECT_NAME_ABBR = "Voltrix-SpMM", "Voltrix"
PROJECT_NAME_FULL_LOWER, PROJECT_NAME_ABBR_LOWER = PROJECT_NAME_FULL.lower(), PROJECT_NAME_ABBR.lower()

_ENV = PROJECT_NAME_ABBR.upper() + "_"


def _env(suffix: str) -> str:
    return _ENV + suffix


# read by the JIT layer (voltrix/jit/compiler.py, jit_kernels/tuner.py), as in the reference
CACHE_DIR_FLAG = _env("CACHE_DIR")                                   # where kernel.<name>.<hash>/ directories live
NVCC_COMPILER_FLAG = _env("NVCC_COMPILER")                           # nvcc to use instead of $CUDA_HOME/bin/nvcc
DEBUG_FLAG = _env("JIT_DEBUG")                                       # trace cache hits / compiles
JIT_PRINT_NVCC_COMMAND_FLAG = _env("JIT_PRINT_NVCC_COMMAND")         # echo every nvcc command line
PTXAS_VERBOSE_FLAG = _env("PTXAS_VERBOSE")                           # -Xptxas -v and print its report
PRINT_AUTOTUNE_FLAG = _env("PRINT_AUTO_TUNE")                        # print every candidate's time while tuning

# ---- new in this implementation ----
FP32_MODE_FLAG = _env("FP32_MODE")                                   # tf32 | split | exact: precision class of fp32 operands
FP32_EXACT_FLAG = _env("FP32_EXACT")                                 # legacy spelling of FP32_MODE=exact ("1")
EXTRA_NVCC_FLAGS_FLAG = _env("EXTRA_NVCC_FLAGS")                     # appended to every JIT compile (part of the cache key)
