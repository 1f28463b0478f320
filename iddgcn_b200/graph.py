"""Graph handle and the reference-facing adjacency objects.

`get_adj_mats(data, num_entities, num_relations)` keeps the signature of the reference's
utils1.get_adj_mats (prediction/utils1.py:420-451) and returns a list of R `Adjacency` objects
that quack like the `tf.SparseTensor(1, N, N)` the reference passes around (`.indices`
`(nnz,3) int64 [0,h,t]`, `.values` `(nnz,) float32`, `adj * mask`), while sharing ONE device
graph (relation-segmented CSR + CSC built by libiddgcn_b200 on the GPU).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import VALS_EXPLICIT, VALS_ONES, VALS_SIGMOID, check, load


class _DevArray:
    """Zero-copy view of a library-owned int32 device array (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        self._owner = owner


def _as_device_triples(data, device) -> torch.Tensor:
    if isinstance(data, torch.Tensor):
        t = data
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(data)).astype(np.int64, copy=False))
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError(f"triples must have shape (M, 3) [head, rel, tail]; got {tuple(t.shape)}")
    return t.contiguous()


class Graph:
    """Owns an `iddgcn_graph_t`: all R relations of one (head, rel, tail) triple set on one GPU."""

    def __init__(self, triples, num_entities: int, num_relations: int, *, device=None,
                 row_block: tuple[int, int] | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("iddgcn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        lib = load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        t = _as_device_triples(triples, self.device)
        lo, hi = (0, int(num_entities)) if row_block is None else (int(row_block[0]), int(row_block[1]))
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            on_dev = 1 if t.is_cuda else 0
            if t.is_cuda and t.device != self.device:
                t = t.to(self.device)
            stream = _lib.current_stream_ptr(self.device)
            check(lib.iddgcn_graph_build(t.data_ptr() if t.numel() else None, t.shape[0], on_dev, int(num_entities),
                                         int(num_relations), lo, hi, stream, C.byref(handle)))
        self._h = handle
        self._lib = lib
        n = C.c_int64(); r = C.c_int(); nr = C.c_int64(); rl = C.c_int64(); nnz = C.c_int64()
        per = (C.c_int64 * int(num_relations))()
        check(lib.iddgcn_graph_info(self._h, C.byref(n), C.byref(r), C.byref(nr), C.byref(rl), C.byref(nnz), per))
        self.N, self.R, self.n_rows, self.row_lo, self.nnz = n.value, r.value, nr.value, rl.value, nnz.value
        self.nnz_per_rel = [int(x) for x in per]
        self.rel_offsets = np.concatenate([[0], np.cumsum(self.nnz_per_rel)]).astype(np.int64)
        self._dev_cache = {}
        self._models = []              # weak refs to models that captured a train step on this graph

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            for ref in getattr(self, "_models", []):
                m = ref()
                if m is not None and getattr(m, "_h", None):
                    self._lib.iddgcn_model_release_graph(m._h, self._h)
            self._lib.iddgcn_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("graph handle was destroyed")
        return self._h

    # -- integer structure ------------------------------------------------------------------
    def export(self) -> dict[str, np.ndarray]:
        """Host copies of the CSR/CSC arrays (for parity tests)."""
        out = {
            "csr_rowptr": np.empty(self.R * self.n_rows + 1, np.int32),
            "csr_col": np.empty(self.nnz, np.int32),
            "csr_row": np.empty(self.nnz, np.int32),
            "csc_rowptr": np.empty(self.R * self.N + 1, np.int32),
            "csc_row": np.empty(self.nnz, np.int32),
            "csc_perm": np.empty(self.nnz, np.int32),
        }
        check(self._lib.iddgcn_graph_export_host(self.handle, *[out[k].ctypes.data for k in
                                                                ("csr_rowptr", "csr_col", "csr_row", "csc_rowptr",
                                                                 "csc_row", "csc_perm")]))
        return out

    def device_arrays(self) -> dict[str, torch.Tensor]:
        """Zero-copy int32 CUDA tensors over the library-owned arrays (valid while the graph lives)."""
        if not self._dev_cache:
            ptrs = [C.c_void_p() for _ in range(6)]
            check(self._lib.iddgcn_graph_device_ptrs(self.handle, *[C.byref(p) for p in ptrs]))
            sizes = [self.R * self.n_rows + 1, self.nnz, self.nnz, self.R * self.N + 1, self.nnz, self.nnz]
            names = ["csr_rowptr", "csr_col", "csr_row", "csc_rowptr", "csc_row", "csc_perm"]
            for name, p, n in zip(names, ptrs, sizes):
                if n == 0:
                    self._dev_cache[name] = torch.empty(0, dtype=torch.int32, device=self.device)
                else:
                    self._dev_cache[name] = torch.as_tensor(_DevArray(p.value, n, self), device=self.device)
        return self._dev_cache

    def relation_indices(self, r: int) -> torch.Tensor:
        """(nnz_r, 3) int64 rows [0, head, tail], row-major sorted — SparseTensor(1,N,N).indices."""
        d = self.device_arrays()
        lo, hi = int(self.rel_offsets[r]), int(self.rel_offsets[r + 1])
        h = d["csr_row"][lo:hi].to(torch.int64) + self.row_lo
        t = d["csr_col"][lo:hi].to(torch.int64)
        return torch.stack([torch.zeros_like(h), h, t], dim=1)


class Adjacency:
    """One relation's adjacency, API-compatible with what the reference scripts touch on a SparseTensor."""

    def __init__(self, graph: Graph, rel: int, values: torch.Tensor | None = None, *, mask_logits=None):
        self.graph = graph
        self.rel = rel
        self._values = values          # None -> implicit ones (or the (0,0)->0.0 placeholder)
        self.mask_logits = mask_logits  # per-edge logits: a_e = sigmoid(logit) fused in-kernel
        self.dense_shape = (1, graph.N, graph.N)

    # reference: empty relation -> indices [[0,0]] values [0.] (utils1.py:427-429), then reshape (1,N,N)
    @property
    def is_placeholder(self) -> bool:
        return self.graph.nnz_per_rel[self.rel] == 0

    @property
    def nnz(self) -> int:
        return self.graph.nnz_per_rel[self.rel]

    @property
    def indices(self) -> torch.Tensor:
        if self.is_placeholder:
            return torch.zeros((1, 3), dtype=torch.int64, device=self.graph.device)
        return self.graph.relation_indices(self.rel)

    @property
    def values(self) -> torch.Tensor:
        if self.mask_logits is not None:
            return torch.sigmoid(self.mask_logits)
        if self._values is not None:
            return self._values
        if self.is_placeholder:
            return torch.zeros(1, dtype=torch.float32, device=self.graph.device)
        return torch.ones(self.nnz, dtype=torch.float32, device=self.graph.device)

    @property
    def is_ones(self) -> bool:
        return self._values is None and self.mask_logits is None

    def with_values(self, values: torch.Tensor) -> "Adjacency":
        return Adjacency(self.graph, self.rel, values)

    def with_mask_logits(self, logits: torch.Tensor) -> "Adjacency":
        """a_e = adj_e * sigmoid(logits[e]) with the sigmoid fused into the SpMM/SDDMM kernels."""
        if not self.is_ones:
            raise ValueError("with_mask_logits needs an unweighted adjacency")
        return Adjacency(self.graph, self.rel, None, mask_logits=logits)

    def __mul__(self, other):
        """`adj * dense` (SparseDenseCwiseMul; GnnExplainer.py:33) or `adj * per_edge_vector`."""
        other = torch.as_tensor(other, device=self.graph.device)
        idx = self.indices
        if other.dim() == 3:
            g = other[idx[:, 0], idx[:, 1], idx[:, 2]]
        elif other.dim() == 2:
            g = other[idx[:, 1], idx[:, 2]]
        elif other.dim() <= 1:
            g = other
        else:
            raise ValueError("unsupported mask rank")
        return Adjacency(self.graph, self.rel, self.values * g)

    __rmul__ = __mul__


def get_adj_mats(data, num_entities, num_relations, *, device=None) -> list[Adjacency]:
    """Drop-in for utils1.get_adj_mats (prediction/utils1.py:420): list of R adjacencies, values 1.0."""
    g = Graph(data, num_entities, num_relations, device=device)
    return [Adjacency(g, r) for r in range(num_relations)]


def collect_edge_values(adj_list: Sequence[Adjacency]):
    """(graph, vals or None, mode) for a list of R adjacencies sharing one graph."""
    g = adj_list[0].graph
    if len(adj_list) != g.R or any(a.graph is not g for a in adj_list) or [a.rel for a in adj_list] != list(range(g.R)):
        raise ValueError("adjacency list must be the R relations of one graph, in order")
    if all(a.is_ones for a in adj_list):
        return g, None, VALS_ONES
    if all(a.mask_logits is not None or a.is_placeholder for a in adj_list):
        parts = [a.mask_logits[: a.nnz] for a in adj_list if not a.is_placeholder]
        vals = torch.cat(parts) if parts else torch.zeros(0, device=g.device)
        return g, vals.to(torch.float32).contiguous(), VALS_SIGMOID
    parts = [a.values[: a.nnz] for a in adj_list if not a.is_placeholder]
    vals = torch.cat(parts) if parts else torch.zeros(0, device=g.device)
    return g, vals.to(torch.float32).contiguous(), VALS_EXPLICIT
