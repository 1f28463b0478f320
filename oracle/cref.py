"""ctypes loader for oracle/spmm_ref.c — ORACLE / CPU BASELINE, test infrastructure only.

Builds oracle/_build/libspmm_ref.so with gcc (-O3 -fopenmp) on first use.  Only tests/ and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "spmm_ref.c"
OUT = HERE / "_build" / "libspmm_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not OUT.exists() or OUT.stat().st_mtime < SRC.stat().st_mtime:
        OUT.parent.mkdir(exist_ok=True)
        # no -march=native: the .so is built in the authoring container and travels to the GPU box
        cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", str(SRC), "-o", str(OUT)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"gcc failed: {res.stderr}")
    return OUT


def load():
    global _lib
    if _lib is None:
        if not OUT.exists():
            build()
        _lib = C.CDLL(str(OUT))
        _lib.spmm_ref_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(load().spmm_ref_max_threads())


def spmm_coo_seq(row, col, val, X, n_out_rows):
    D = X.shape[1]
    out = np.zeros((n_out_rows, D), np.float32)
    load().spmm_coo_seq(_p(row), _p(col), _p(val), C.c_int64(len(row)), _p(X), C.c_int64(D), _p(out))
    return out


def spmm_csr(ptr, idx, val, X, out=None, threads=0):
    D = X.shape[1]
    nseg = len(ptr) - 1
    if out is None:
        out = np.empty((nseg, D), np.float32)
    load().spmm_csr_omp(_p(ptr), _p(idx), _p(val), C.c_int64(nseg), _p(X), C.c_int64(D), _p(out), C.c_int(threads))
    return out


def spmm_csr_acc64(ptr, idx, val, X, out=None, threads=0):
    """spmm_csr with double accumulators (rounded to float32 at the end) — the accuracy reference for hub rows."""
    D = X.shape[1]
    assert D <= 1024
    nseg = len(ptr) - 1
    if out is None:
        out = np.empty((nseg, D), np.float32)
    load().spmm_csr_omp_acc64(_p(ptr), _p(idx), _p(val), C.c_int64(nseg), _p(X), C.c_int64(D), _p(out),
                              C.c_int(threads))
    return out


def spmm_csc_bwd(cscptr, cscrow, perm, val, R, N, n_rows, dS, out=None, threads=0):
    D = dS.shape[-1]
    if out is None:
        out = np.empty((N, D), np.float32)
    load().spmm_csc_bwd_omp(_p(cscptr), _p(cscrow), _p(perm), _p(val), C.c_int(R), C.c_int64(N), C.c_int64(n_rows),
                            _p(dS), C.c_int64(D), _p(out), C.c_int(threads))
    return out
