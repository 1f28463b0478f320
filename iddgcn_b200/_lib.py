"""ctypes binding of libiddgcn_b200.so (include/iddgcn.h).

There is no CPU fallback and no alternative backend: if the shared library is missing it is
built in place with nvcc; if that is impossible, importing any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import build as _build

_LIB = None

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
vp = C.c_void_p

VALS_ONES, VALS_EXPLICIT, VALS_SIGMOID = 0, 1, 2

_SIGNATURES = {
    # name: (restype, [argtypes])
    "iddgcn_last_error": (C.c_char_p, []),
    "iddgcn_version": (C.c_char_p, []),
    "iddgcn_launch_count": (C.c_int64, []),
    "iddgcn_pool_trim": (C.c_int, []),
    "iddgcn_graph_build": (C.c_int, [vp, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int64, vp,
                                     C.POINTER(vp)]),
    "iddgcn_graph_destroy": (C.c_int, [vp]),
    "iddgcn_graph_info": (C.c_int, [vp, c_i64p, C.POINTER(C.c_int), c_i64p, c_i64p, c_i64p, c_i64p]),
    "iddgcn_graph_export_host": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "iddgcn_graph_device_ptrs": (C.c_int, [vp] + [C.POINTER(vp)] * 6),
    "iddgcn_spmm_fwd": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int, vp, vp]),
    "iddgcn_spmm_bwd_x": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int, vp, C.c_int, vp]),
    "iddgcn_spmm_bwd_vals": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int, vp, vp]),
    "iddgcn_spmm_fwd_rows": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int, vp, C.c_int64, C.c_int64, vp]),
    "iddgcn_spmm_bwd_x_rows": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int, vp, C.c_int64, C.c_int64, C.POINTER(vp),
                                         C.c_int, vp]),
    "iddgcn_spmm_prepare": (C.c_int, [vp, C.c_int64, vp]),
    "iddgcn_peer_signal": (C.c_int, [C.POINTER(vp), C.c_int, C.c_uint32, vp]),
    "iddgcn_peer_wait": (C.c_int, [C.POINTER(vp), C.c_int, C.c_uint32, C.c_int, vp, vp]),
    "iddgcn_copy_async": (C.c_int, [vp, vp, C.c_int64, vp]),
    "iddgcn_copy2d_async": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_int64, C.c_int64, vp]),
    "iddgcn_peer_push": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp]),
    "iddgcn_edge_value_sums": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "iddgcn_spmm_fwd_bwd_host": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp]),
    "iddgcn_spmm_fwd_dl": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "iddgcn_spmm_bwd_x_dl": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_int, vp]),
    "iddgcn_layer_fwd_dl": (C.c_int, [vp] * 9),
    "iddgcn_layer_bwd_dl": (C.c_int, [vp] * 10 + [C.c_int] + [vp] * 4),
    "iddgcn_model_forward_dl": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp]),
    "iddgcn_plan_create": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int, C.c_int64, C.c_int, vp, C.POINTER(vp)]),
    "iddgcn_plan_destroy": (C.c_int, [vp]),
    "iddgcn_relation_transform": (C.c_int, [vp, vp, C.c_int, C.c_int64, C.c_int64, vp, vp]),
    "iddgcn_relation_transform_bwd": (C.c_int, [vp, vp, vp, C.c_int, C.c_int64, C.c_int64, vp, vp, C.c_int, vp]),
    "iddgcn_dense_transform": (C.c_int, [vp, vp, C.c_int64, C.c_int64, C.c_int, vp, C.c_int, C.c_int, vp]),
    "iddgcn_dense_transform_grad": (C.c_int, [vp, vp, C.c_int64, C.c_int64, vp, C.c_int, C.c_int, vp]),
    "iddgcn_layer_fwd": (C.c_int, [vp] * 8 + [C.c_int64, vp, vp, vp, vp]),
    "iddgcn_layer_bwd": (C.c_int, [vp] * 7 + [C.c_int64, vp, vp, vp, C.c_int, vp, vp, vp, vp]),
    "iddgcn_segment_scatter": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int, vp]),
    "iddgcn_distmult_fwd": (C.c_int, [vp, vp, vp, C.c_int64, vp, vp]),
    "iddgcn_distmult_bwd": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "iddgcn_bce_loss_grad": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_float, vp, vp, vp]),
    "iddgcn_model_create": (C.c_int, [C.c_int64, C.c_int, C.c_int64, C.c_int, C.POINTER(vp)]),
    "iddgcn_model_destroy": (C.c_int, [vp]),
    "iddgcn_model_num_params": (C.c_int64, [vp]),
    "iddgcn_model_param_layout": (C.c_int, [vp, c_i64p, c_i64p]),
    "iddgcn_model_buffers": (C.c_int, [vp] + [C.POINTER(vp)] * 4),
    "iddgcn_model_release_plan": (C.c_int, [vp, vp]),
    "iddgcn_model_release_graph": (C.c_int, [vp, vp]),
    "iddgcn_model_profile": (C.c_int, [vp, C.c_int]),
    "iddgcn_model_profile_read": (C.c_int, [vp, vp, C.c_int]),
    "iddgcn_model_forward": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp]),
    "iddgcn_model_train_step": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64,
                                          C.c_int, vp, vp, vp]),
    "iddgcn_model_train_loop": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64,
                                          C.c_int64, vp, vp]),
    "iddgcn_model_backward": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, vp]),
    "iddgcn_subgraph_triples": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int64, c_i64p, vp]),
    "iddgcn_block_relation_sums": (C.c_int, [vp, vp, C.c_int, C.c_int64, C.c_int64, vp, vp]),
    "iddgcn_explain_loss": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int64, vp, vp, vp, vp]),
    "iddgcn_structure_grad": (C.c_int, [vp, vp, C.c_int, vp, C.c_int64, C.c_int64, vp, vp]),
    "iddgcn_mc_store": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp]),
    "iddgcn_mc_reduce_load": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp]),
    "iddgcn_adam_step": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64,
                                   vp]),
}


class IddgcnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libiddgcn_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> Path:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building if necessary) the shared library.  Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists() or os.environ.get("IDDGCN_REBUILD") == "1":
        _build.build()
    if not path.exists():
        raise ImportError(f"{path} is missing and could not be built; iddgcn_b200 has no CPU fallback")
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue  # tests/test_abi.py checks the export list against include/iddgcn.h
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise IddgcnError(rc, load().iddgcn_last_error().decode())


def ptr(t) -> int | None:
    """Device (or host) address of a torch tensor; None passes NULL."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(load().iddgcn_launch_count())
