"""GPU parity of the dense transform C (+)= A·W / A·Wᵀ (tf.matmul(x, kernel), prediction/IDDGCN.py:62-63,76-77) through
the C ABI, both implementations: the fp32 SIMT kernels and the tensor-core kernel (tcgen05.mma kind::tf32 with the 3xTF32
split, TMA-staged operands, TMEM accumulators).

Tolerances, relative to Σ_k |a_ik|·|w_kj| (the scale rounding errors live on), against an fp64 product:
  fp32 SIMT            1e-6   (K ≤ 128 sequential fp32 FMAs)
  tensor cores, 3xTF32 2e-6   (dropped lo·lo term + truncation of the lo operands, ~2^-21 per product; plain TF32 would
                               be ~5e-4 and fail this bar by two orders of magnitude)
"""
import ctypes as C

import numpy as np
import pytest
import torch

from iddgcn_b200 import _lib

pytestmark = pytest.mark.gpu

SIMT_TOL = 1e-6
TC_TOL = 2e-6


def _transform(A, W, trans_w, Cinit, path):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    a = torch.as_tensor(A, device=dev).contiguous()
    w = torch.as_tensor(W, device=dev).contiguous()
    c = torch.as_tensor(Cinit, device=dev).clone() if Cinit is not None else torch.full(
        (A.shape[0], W.shape[0]), float("nan"), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.iddgcn_dense_transform(_lib.ptr(a), _lib.ptr(w), A.shape[0], W.shape[0], int(trans_w), _lib.ptr(c),
                                          int(Cinit is not None), path, C.c_void_p(st)))
    torch.cuda.synchronize()
    return c.cpu().numpy()


def _case(M, D, seed, wide_range=False):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, D)).astype(np.float32)
    if wide_range:                                        # magnitudes over 12 binades: the split must be per element
        A *= np.exp2(rng.integers(-6, 7, (M, D))).astype(np.float32)
    W = (rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32)
    C0 = rng.standard_normal((M, D)).astype(np.float32)
    return A, W, C0


def _scaled_err(got, A, W, trans_w, C0):
    Wm = W.T if trans_w else W
    ref = A.astype(np.float64) @ Wm.astype(np.float64) + (C0.astype(np.float64) if C0 is not None else 0.0)
    scale = np.abs(A).astype(np.float64) @ np.abs(Wm).astype(np.float64) + (np.abs(C0) if C0 is not None else 0.0)
    return float(np.max(np.abs(got - ref) / scale))


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("M", [1, 100, 128, 129, 1000, 40000])
@pytest.mark.parametrize("trans_w", [0, 1])
def test_tensor_core_transform_matches_fp64(D, M, trans_w):
    A, W, C0 = _case(M, D, seed=M + D + trans_w, wide_range=(M == 1000))
    for acc in (None, C0):
        got = _transform(A, W, trans_w, acc, path=2)
        assert np.isfinite(got).all()
        assert _scaled_err(got, A, W, trans_w, acc) < TC_TOL
        simt = _transform(A, W, trans_w, acc, path=1)
        assert _scaled_err(simt, A, W, trans_w, acc) < SIMT_TOL


def test_three_term_split_is_really_used():
    """Plain TF32 keeps 10 mantissa bits: x = 1 + 2^-12 would lose its low bit and the product with the identity would
    come back as 1.  With the hi/lo split the result is exact."""
    D, M = 128, 256
    A = np.full((M, D), 1.0 + 2.0 ** -12, np.float32)
    A[::2] = 3.0 - 2.0 ** -20
    W = np.eye(D, dtype=np.float32)
    got = _transform(A, W, 0, None, path=2)
    assert np.array_equal(got, A)
    Wl = np.eye(D, dtype=np.float32) * np.float32(1.0 + 2.0 ** -13)          # the low part of W matters too
    got = _transform(np.ones((M, D), np.float32), Wl, 0, None, path=2)
    assert np.array_equal(got, np.ones((M, D), np.float32) * np.float32(1.0 + 2.0 ** -13))


def test_automatic_choice_and_switch(monkeypatch):
    """path 0: SIMT below 32768 rows, tensor cores from there on (bit-identical to the forced paths);
    IDDGCN_GEMM_TF32X3=0 keeps everything on the SIMT kernels."""
    A, W, _ = _case(40000, 128, seed=5)
    assert np.array_equal(_transform(A, W, 0, None, 0), _transform(A, W, 0, None, 2))
    assert np.array_equal(_transform(A[:1000], W, 0, None, 0), _transform(A[:1000], W, 0, None, 1))
    monkeypatch.setenv("IDDGCN_GEMM_TF32X3", "0")
    assert np.array_equal(_transform(A, W, 0, None, 0), _transform(A, W, 0, None, 1))


def test_tensor_core_path_rejects_other_widths():
    A, W, _ = _case(64, 32, seed=1)
    with pytest.raises(_lib.IddgcnError):
        _transform(A, W, 0, None, path=2)
    assert np.isfinite(_transform(A, W, 0, None, path=0)).all()                  # automatic falls back to SIMT


def test_misaligned_output_falls_back_or_is_refused():
    """The tensor-core epilogue stores 256 bits at a time: a C that is only 16-byte aligned is refused on the forced
    path and silently served by the SIMT kernels on the automatic one."""
    lib = _lib.load()
    dev = torch.device("cuda:0")
    M, D = 40000, 64
    A, W, _ = _case(M, D, seed=2)
    a, w = torch.as_tensor(A, device=dev), torch.as_tensor(W, device=dev)
    buf = torch.zeros(M * D + 8, device=dev)
    c = buf[4:4 + M * D]                                                     # 16-byte aligned, not 32
    assert c.data_ptr() % 32 == 16
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with pytest.raises(_lib.IddgcnError):
        _lib.check(lib.iddgcn_dense_transform(_lib.ptr(a), _lib.ptr(w), M, D, 0, c.data_ptr(), 0, 2, st))
    _lib.check(lib.iddgcn_dense_transform(_lib.ptr(a), _lib.ptr(w), M, D, 0, c.data_ptr(), 0, 0, st))
    torch.cuda.synchronize()
    assert _scaled_err(c.view(M, D).cpu().numpy(), A, W, 0, None) < SIMT_TOL


def test_many_tiles_per_cta_and_determinism():
    """More tiles than SMs (the persistent loop wraps the stage ring and both accumulators many times), run twice."""
    A, W, C0 = _case(148 * 128 * 5 + 77, 64, seed=9)
    a = _transform(A, W, 1, C0, path=2)
    b = _transform(A, W, 1, C0, path=2)
    assert np.array_equal(a, b)
    assert _scaled_err(a, A, W, 1, C0) < TC_TOL


# ---- weight gradient dW (+)= Aᵀ·dC ------------------------------------------------------------------------------------
def _grad(A, dC, W0, path):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    a = torch.as_tensor(A, device=dev).contiguous()
    d = torch.as_tensor(dC, device=dev).contiguous()
    D = A.shape[1]
    w = torch.as_tensor(W0, device=dev).clone() if W0 is not None else torch.full((D, D), float("nan"), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.iddgcn_dense_transform_grad(_lib.ptr(a), _lib.ptr(d), A.shape[0], D, _lib.ptr(w),
                                               int(W0 is not None), path, C.c_void_p(st)))
    torch.cuda.synchronize()
    return w.cpu().numpy()


def _grad_err(got, A, dC, W0):
    ref = A.astype(np.float64).T @ dC.astype(np.float64) + (W0.astype(np.float64) if W0 is not None else 0.0)
    scale = np.abs(A).astype(np.float64).T @ np.abs(dC).astype(np.float64) + (np.abs(W0) if W0 is not None else 0.0)
    return float(np.max(np.abs(got - ref) / scale))


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("M", [1, 31, 32, 33, 255, 256, 257, 5000, 300000])
def test_tensor_core_weight_gradient_matches_fp64(D, M):
    """Row counts around the 32-row k-block and the 256-row TMEM flush, fewer k-blocks than SMs, and a long reduction."""
    rng = np.random.default_rng(M + D)
    A = rng.standard_normal((M, D)).astype(np.float32)
    dC = rng.standard_normal((M, D)).astype(np.float32)
    if M == 5000:
        A *= np.exp2(rng.integers(-6, 7, (M, D))).astype(np.float32)
    W0 = rng.standard_normal((D, D)).astype(np.float32)
    for acc in (None, W0):
        got = _grad(A, dC, acc, path=2)
        assert np.isfinite(got).all()
        assert _grad_err(got, A, dC, acc) < TC_TOL
        assert _grad_err(_grad(A, dC, acc, path=1), A, dC, acc) < SIMT_TOL


def test_weight_gradient_is_exact_on_split_sensitive_values_and_deterministic():
    """Aᵀ·dC with dC = [I; 0; ...] picks single rows of A: exact only if hi and lo both arrive, transposed correctly."""
    D, M = 128, 4096
    rng = np.random.default_rng(0)
    A = (1.0 + rng.integers(0, 2 ** 20, (M, D)) * 2.0 ** -20).astype(np.float32)        # full 21-bit mantissas
    dC = np.zeros((M, D), np.float32)
    rows = rng.permutation(M)[:D]
    dC[rows, np.arange(D)] = 1.0                                                          # dW[i, j] = A[rows[j], i]
    got = _grad(A, dC, None, path=2)
    assert np.array_equal(got, A[rows].T)
    assert np.array_equal(got, _grad(A, dC, None, path=2))


def test_weight_gradient_automatic_choice(monkeypatch):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((40000, 128)).astype(np.float32)
    dC = rng.standard_normal((40000, 128)).astype(np.float32)
    assert np.array_equal(_grad(A, dC, None, 0), _grad(A, dC, None, 2))
    assert np.array_equal(_grad(A[:1000], dC[:1000], None, 0), _grad(A[:1000], dC[:1000], None, 1))
    monkeypatch.setenv("IDDGCN_GEMM_TF32X3", "0")
    assert np.array_equal(_grad(A, dC, None, 0), _grad(A, dC, None, 1))
