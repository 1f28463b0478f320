"""Triple plans: the (head, rel, tail) index set of one model call, kept on device with the by-head / by-tail /
by-relation orderings that make every gather's backward an atomics-free segment sum (iddgcn_plan_t)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, load


def _as_i64(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)).astype(np.int64, copy=False))
    t = t.reshape(-1)
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


class Plan:
    def __init__(self, head, rel, tail, num_entities: int, num_relations: int, *, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("iddgcn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h, r, t = (_as_i64(x, self.device) for x in (head, rel, tail))
        if not (h.numel() == r.numel() == t.numel()):
            raise ValueError("head, rel, tail must have the same length")
        on_dev = h.is_cuda and r.is_cuda and t.is_cuda
        if not on_dev:
            h, r, t = h.cpu(), r.cpu(), t.cpu()
        self.T = h.numel()
        self.N, self.R = int(num_entities), int(num_relations)
        handle = C.c_void_p()
        self._lib = load()
        with torch.cuda.device(self.device):
            check(self._lib.iddgcn_plan_create(h.data_ptr() if self.T else None, r.data_ptr() if self.T else None,
                                               t.data_ptr() if self.T else None, self.T, int(on_dev), self.N, self.R,
                                               _lib.current_stream_ptr(self.device), C.byref(handle)))
        self._h = handle
        self._models = []

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("plan handle was destroyed")
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            for ref in self._models:
                m = ref()
                if m is not None and m._h:
                    self._lib.iddgcn_model_release_plan(m._h, self._h)
            self._lib.iddgcn_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
