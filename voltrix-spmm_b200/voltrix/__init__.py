"""voltrix -- B200-native drop-in for the Voltrix-SpMM hot path (reference: voltrix/__init__.py:1-3)."""
from .project import *  # noqa: F401,F403
from .jit_kernels import *  # noqa: F401,F403
from .spmm import *  # noqa: F401,F403
from .spmm import BLK_H, BLK_W, csr_preprocess, spmm, SpmmPlan, HostStreamedSpMM, gcn_norm, spmm_gcn, spmm_weighted, save_preprocessed, load_preprocessed, EdgeWeights, edge_weights, reschedule, tune_routing, ROUTING_CANDIDATES  # noqa: F401  (`spmm` the function shadows the sub-package, as in the reference)
from . import autograd, graphs, jit, jit_kernels, project, reorder, utils  # noqa: F401
from .autograd import SparseAdj, csr_transpose, spmm_autograd  # noqa: F401
