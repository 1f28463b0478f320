"""iddgcn_b200 — B200-native (sm_100a) implementation of IDDGCN's relational graph-convolution hot path.

Python here is the host side of a C ABI (include/iddgcn.h, iddgcn_b200/csrc/*.cu): it mirrors the reference's
layer / model / get_adj_mats interface and owns device memory and streams through torch.  There is no CPU
fallback and no alternative backend.
"""
from .graph import Adjacency, Graph, get_adj_mats  # noqa: F401
from .layers import DistMult, IDDGCN_Layer  # noqa: F401
from .model import Adam, IDDGCN_Model, get_IDDGCN_Model  # noqa: F401
from .plan import Plan  # noqa: F401

__all__ = ["Adjacency", "Graph", "get_adj_mats", "IDDGCN_Layer", "DistMult", "IDDGCN_Model", "get_IDDGCN_Model",
           "Adam", "Plan"]
