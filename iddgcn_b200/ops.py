"""torch.autograd bindings of the C-ABI operators (PyTorch is plumbing: memory, streams, autograd tape).

Every function here launches hand-written sm_100a kernels through ctypes; nothing falls back to
torch ops or to the CPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import VALS_EXPLICIT, VALS_ONES, VALS_SIGMOID, check, load, ptr
from .graph import Graph


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: iddgcn_b200 has no CPU path")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def spmm_fwd(graph: Graph, X: torch.Tensor, vals=None, mode: int = VALS_ONES, out=None) -> torch.Tensor:
    """S[r] = A_r · X for all relations (tf.sparse.sparse_dense_matmul, IDDGCN.py:70). S: (R, n_rows, D)."""
    X = _f32c(X, "X")
    D = X.shape[1]
    if X.shape[0] != graph.N:
        raise ValueError(f"X has {X.shape[0]} rows, graph has N={graph.N}")
    S = out if out is not None else torch.empty((graph.R, graph.n_rows, D), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):          # launches go to the current device: make it the tensors' device
        check(load().iddgcn_spmm_fwd(graph.handle, ptr(X), D, ptr(vals), mode, ptr(S), _stream(X)))
    return S


def spmm_bwd_x(graph: Graph, dS: torch.Tensor, vals=None, mode: int = VALS_ONES, out=None,
               accumulate: bool = False) -> torch.Tensor:
    """dX = Σ_r A_rᵀ · dS_r (adjoint SpMM, relations fused, no atomics). dX: (N, D)."""
    dS = _f32c(dS, "dS")
    D = dS.shape[-1]
    dX = out if out is not None else torch.empty((graph.N, D), dtype=torch.float32, device=dS.device)
    with torch.cuda.device(dS.device):
        check(load().iddgcn_spmm_bwd_x(graph.handle, ptr(dS), D, ptr(vals), mode, ptr(dX), int(accumulate),
                                       _stream(dS)))
    return dX


def spmm_fwd_rows(graph: Graph, X: torch.Tensor, S: torch.Tensor, row_begin: int, row_end: int, vals=None,
                  mode: int = VALS_ONES) -> None:
    """Rows [row_begin, row_end) of the flattened (R·n_rows, D) output only (pipelined callers); S is the FULL buffer."""
    X = _f32c(X, "X")
    D = X.shape[1]
    Sf = S.view(-1, D)
    with torch.cuda.device(X.device):
        check(load().iddgcn_spmm_fwd_rows(graph.handle, ptr(X), D, ptr(vals), mode, Sf[row_begin:].data_ptr(),
                                          row_begin, row_end, _stream(X)))


def spmm_bwd_x_rows(graph: Graph, dS: torch.Tensor, dX_rows: torch.Tensor, node_begin: int, node_end: int,
                    addends=(), vals=None, mode: int = VALS_ONES) -> None:
    """dX rows of source nodes [node_begin, node_end) into dX_rows (first row = node_begin; may be peer memory), plus
    the addends (tensors or raw addresses, each at the first row of the range) summed in the kernel epilogue."""
    dS = _f32c(dS, "dS")
    D = dS.shape[-1]
    addrs = [a if isinstance(a, int) else a.data_ptr() for a in addends]
    arr = (C.c_void_p * max(len(addrs), 1))(*addrs) if addrs else None
    dst = dX_rows if isinstance(dX_rows, int) else dX_rows.data_ptr()
    with torch.cuda.device(dS.device):
        check(load().iddgcn_spmm_bwd_x_rows(graph.handle, ptr(dS), D, ptr(vals), mode, dst, node_begin, node_end, arr,
                                            len(addrs), _stream(dS)))


def spmm_bwd_vals(graph: Graph, dS: torch.Tensor, X: torch.Tensor, vals=None, mode: int = VALS_ONES,
                  out=None) -> torch.Tensor:
    """∂L/∂vals[e] = <dS[row_e], X[col_e]> (·σ'(v) in SIGMOID mode) — SDDMM on the CSR pattern."""
    dS = _f32c(dS, "dS")
    X = _f32c(X, "X")
    D = X.shape[1]
    dv = out if out is not None else torch.empty(graph.nnz, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(load().iddgcn_spmm_bwd_vals(graph.handle, ptr(dS), ptr(X), D, ptr(vals), mode, ptr(dv), _stream(X)))
    return dv


def edge_value_sums(graph: Graph, vals=None, mode: int = VALS_ONES) -> torch.Tensor:
    out = torch.empty(graph.R, dtype=torch.float32, device=graph.device)
    with torch.cuda.device(graph.device):
        check(load().iddgcn_edge_value_sums(graph.handle, ptr(vals), mode, ptr(out),
                                            torch.cuda.current_stream(graph.device).cuda_stream))
    return out


def _dl_ptr(capsule) -> int:
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.pythonapi.PyCapsule_GetPointer(capsule, b"dltensor")


def spmm_fwd_dlpack(graph: Graph, X, S, vals=None, mode: int = VALS_ONES) -> None:
    """Same as spmm_fwd but every tensor crosses the boundary as a DLPack capsule (any producer)."""
    caps = [X.__dlpack__(), S.__dlpack__(), vals.__dlpack__() if vals is not None else None]
    stream = torch.cuda.current_stream(graph.device).cuda_stream
    check(load().iddgcn_spmm_fwd_dl(graph.handle, _dl_ptr(caps[0]), _dl_ptr(caps[2]) if caps[2] is not None else None,
                                    mode, _dl_ptr(caps[1]), stream))


def spmm_bwd_x_dlpack(graph: Graph, dS, dX, vals=None, mode: int = VALS_ONES, accumulate: bool = False) -> None:
    caps = [dS.__dlpack__(), dX.__dlpack__(), vals.__dlpack__() if vals is not None else None]
    stream = torch.cuda.current_stream(graph.device).cuda_stream
    check(load().iddgcn_spmm_bwd_x_dl(graph.handle, _dl_ptr(caps[0]),
                                      _dl_ptr(caps[2]) if caps[2] is not None else None, mode, _dl_ptr(caps[1]),
                                      int(accumulate), stream))


def layer_fwd_dlpack(plan, HT_in, Z, W_self, W_alpha, b_alpha, HT_out, alpha_out) -> None:
    """iddgcn_layer_fwd with every tensor crossing the boundary as a DLPack capsule (IDDGCN.py:60-79)."""
    ts = [HT_in, Z, W_self, W_alpha, b_alpha, HT_out, alpha_out]
    caps = [t.__dlpack__() for t in ts]
    check(load().iddgcn_layer_fwd_dl(plan.handle, *[_dl_ptr(c) for c in caps],
                                     torch.cuda.current_stream(plan.device).cuda_stream))


def layer_bwd_dlpack(plan, HT_in, HT_out, alpha, Z, W_self, W_alpha, dHT_out, dHT_in, dZ, dW_self, dW_alpha, db_alpha,
                     accumulate_dZ: bool = False) -> None:
    """iddgcn_layer_bwd through DLPack capsules; weight gradients are accumulated into dW_self / dW_alpha / db_alpha."""
    a = [t.__dlpack__() for t in (HT_in, HT_out, alpha, Z, W_self, W_alpha, dHT_out, dHT_in, dZ)]
    b = [t.__dlpack__() for t in (dW_self, dW_alpha, db_alpha)]
    check(load().iddgcn_layer_bwd_dl(plan.handle, *[_dl_ptr(c) for c in a], int(accumulate_dZ),
                                     *[_dl_ptr(c) for c in b], torch.cuda.current_stream(plan.device).cuda_stream))


def model_forward_dlpack(model, graph: Graph, plan, y, vals=None, mode: int = VALS_ONES) -> None:
    """iddgcn_model_forward through DLPack capsules: y [T] <- model(triples of `plan` | graph, optional edge values)."""
    cy = y.__dlpack__()
    cv = vals.__dlpack__() if vals is not None else None
    check(load().iddgcn_model_forward_dl(model._h, graph.handle, plan.handle, _dl_ptr(cv) if cv is not None else None,
                                         mode, _dl_ptr(cy), torch.cuda.current_stream(graph.device).cuda_stream))


def spmm_fwd_bwd_host(graph: Graph, X_host: torch.Tensor, dS_host: torch.Tensor, S_host=None, dX_host=None):
    """End-to-end call with HOST buffers (pinned recommended): upload, fwd, bwd, download."""
    for t, n in ((X_host, "X_host"), (dS_host, "dS_host")):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{n} must be a contiguous float32 host tensor")
    D = X_host.shape[1]
    with torch.cuda.device(graph.device):
        check(load().iddgcn_spmm_fwd_bwd_host(graph.handle, ptr(X_host), ptr(dS_host), D, ptr(S_host), ptr(dX_host)))


class _RelationalSpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, vals, graph, mode):
        Xc = _f32c(X, "X")
        vc = _f32c(vals, "vals") if vals is not None else None
        ctx.graph, ctx.mode = graph, mode
        ctx.save_for_backward(Xc, vc)
        return spmm_fwd(graph, Xc, vc, mode)

    @staticmethod
    def backward(ctx, dS):
        X, vals = ctx.saved_tensors
        dS = _f32c(dS, "dS")
        dX = dvals = None
        if ctx.needs_input_grad[0]:
            dX = spmm_bwd_x(ctx.graph, dS, vals, ctx.mode)
        if vals is not None and ctx.needs_input_grad[1]:
            dvals = spmm_bwd_vals(ctx.graph, dS, X, vals, ctx.mode)
        return dX, dvals, None, None


def relational_spmm(graph: Graph, X: torch.Tensor, vals=None, mode: int = VALS_ONES) -> torch.Tensor:
    """Differentiable S = A·X over all relations; gradients w.r.t. X and (if given) the edge values."""
    return _RelationalSpMM.apply(X, vals, graph, mode)
