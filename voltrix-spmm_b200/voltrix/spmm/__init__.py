from .spmm import BLK_H, BLK_W
from .spmm import (
    csr_preprocess,
    spmm,
    SpmmPlan,
    HostStreamedSpMM,
    gcn_norm,
    spmm_weighted,
    save_preprocessed,
    load_preprocessed,
    spmm_gcn,
    EdgeWeights,
    edge_weights,
    reschedule,
    tune_routing,
    ROUTING_CANDIDATES,
)
