"""GPU parity of the gather-backward segment sums (iddgcn_segment_scatter, SURVEY App. A eq. 10 and 12): the flat-stream
kernel used for large plans against an fp64 NumPy scatter-add and against the warp-per-node kernel it replaces there."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _scatter(p, dP, alpha, out, accumulate):
    from iddgcn_b200._lib import check, load, ptr
    check(load().iddgcn_segment_scatter(p.handle, ptr(dP), ptr(alpha), dP.shape[-1], ptr(out), int(accumulate),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))


@pytest.mark.parametrize("D,R", [(16, 1), (64, 2), (128, 2), (128, 4), (32, 3)])
@pytest.mark.parametrize("gated", [True, False])
def test_segment_scatter_stream_kernel(D, R, gated, monkeypatch):
    from iddgcn_b200.plan import Plan
    rng = np.random.default_rng(D + R)
    N, T = 3000, 20000
    head = rng.integers(0, N, T); tail = rng.integers(0, N, T); rel = rng.integers(0, R, T)
    head[:4000] = 17                                   # a hub node with thousands of triples (Kahan sums)
    tail[(tail % 5) == 2] = 1                          # runs of nodes that own nothing on the tail side
    dP = rng.standard_normal((2, T, D)).astype(np.float32)
    alpha = rng.standard_normal((T, R)).astype(np.float32) if gated else None
    planes = R if gated else 1
    ref = np.zeros((planes, N, D))
    for r in range(planes):
        w = 1.0 / (1.0 + np.exp(-alpha[:, r].astype(np.float64))) if gated else np.ones(T)
        np.add.at(ref[r], head, w[:, None] * dP[0])
        np.add.at(ref[r], tail, w[:, None] * dP[1])
    plan = Plan(head, rel, tail, N, R)
    dPd = torch.from_numpy(dP).cuda()
    ad = torch.from_numpy(alpha).cuda() if gated else None
    outs = {}
    for mode in ("1", "0"):                            # flat-stream kernel, then the warp-per-node kernel
        monkeypatch.setenv("IDDGCN_SEG_STREAM", mode)
        out = torch.full((planes, N, D), -3.0, device="cuda")
        _scatter(plan, dPd, ad, out, False)
        assert rel_err(out.cpu().numpy(), ref) < 1e-6, mode
        base = torch.ones((planes, N, D), device="cuda")
        _scatter(plan, dPd, ad, base, True)
        assert rel_err(base.cpu().numpy(), ref + 1.0) < 1e-6, mode
        outs[mode] = out
    assert rel_err(outs["1"].cpu().numpy(), outs["0"].cpu().numpy()) < 1e-6
