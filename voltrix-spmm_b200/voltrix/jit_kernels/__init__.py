from .bmat_swizzle import hmat_packed_swizzle_kernel
from .hmat_gem import hmat_gen_kernel
from .spmm import spmm_kernel
from .spmm_csr import spmm_csr_weighted_kernel
from .preprocess import preprocess_kernel
from .tiles import (csr_window_sort_kernel, csr_tiles_scatter_kernel, preprocess_workspace_bytes,
                    schedule_build_kernel, schedule_sort_kernel, schedule_sizes)
from .value_tiles import value_tiles_kernel
from .tuner import jit_tuner
