"""N>1 host logic on the CPU: world_size-2 gloo run of the row-partitioned SpMM exchange (iddgcn_b200/dist.py).

The CUDA kernels cannot run here, so the rank-local compute is injected from the oracle (tests are allowed to use
it); what is under test is the partition arithmetic, the panel layout, and the all-gather / reduce-scatter plumbing:
the concatenated rank outputs must equal the single-process result.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_triples
from oracle import iddgcn_oracle as O


def _worker(rank, world, port, N, R, D, panels, tri, X, dS, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iddgcn_b200.dist import PartitionedSpMM, row_block
        lo, hi, block = row_block(rank, world, N)
        adj = O.get_adj_mats(tri, N, R)
        local = [O.SparseCOO(a.indices[(a.indices[:, 1] >= lo) & (a.indices[:, 1] < hi) & (a.values != 0)] -
                             np.array([0, lo, 0]), a.values[(a.indices[:, 1] >= lo) & (a.indices[:, 1] < hi) &
                                                            (a.values != 0)], a.dense_shape) for a in adj]
        g = SimpleNamespace(R=R, N=N, n_rows=hi - lo, device=torch.device("cpu"))

        def fwd(g_, Xc, out):       # S_local = A_local · X_full   (rows are local, columns global)
            Xn = Xc.numpy().astype(np.float64)
            S = np.zeros((R, N, Xn.shape[1]))
            for r, a in enumerate(local):
                np.add.at(S[r], a.indices[:, 1], Xn[a.indices[:, 2]])
            out.copy_(torch.from_numpy(S[:, : g_.n_rows].astype(np.float32)))

        def bwd(g_, dSc, out):      # dX_partial = Σ_r A_local,rᵀ · dS_local over ALL source nodes
            dSn = dSc.numpy().astype(np.float64)
            dX = np.zeros((N, dSn.shape[-1]))
            for r, a in enumerate(local):
                np.add.at(dX, a.indices[:, 2], dSn[r][a.indices[:, 1]])
            out.copy_(torch.from_numpy(dX.astype(np.float32)))

        eng = PartitionedSpMM(g, D, world, rank, block, panels=panels, fwd_fn=fwd, bwd_fn=bwd,
                              device=torch.device("cpu"))
        eng.step(torch.from_numpy(X[lo:hi]), torch.from_numpy(dS[:, lo:hi]))
        S_local, dX_local = eng.outputs()
        np.save(os.path.join(out_dir, f"S_{rank}.npy"), S_local.numpy())
        np.save(os.path.join(out_dir, f"dX_{rank}.npy"), dX_local.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,panels", [(101, 4), (64, 2)])
def test_partitioned_spmm_world2_gloo(tmp_path, N, panels):
    R, D, world = 3, 16, 2
    tri = random_triples(N, R, 900, seed=N, heavy_row=2)
    rng = np.random.default_rng(0)
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    port = 29600 + (os.getpid() % 200) + panels
    mp.spawn(_worker, args=(world, port, N, R, D, panels, tri, X, dS, str(tmp_path)), nprocs=world, join=True)
    adj = O.get_adj_mats(tri, N, R)
    S_ref = O.spmm_fwd(adj, X.astype(np.float64))
    dX_ref = O.spmm_bwd_x(adj, dS.astype(np.float64))
    S = np.concatenate([np.load(tmp_path / f"S_{r}.npy") for r in range(world)], axis=1)
    dX = np.concatenate([np.load(tmp_path / f"dX_{r}.npy") for r in range(world)], axis=0)
    assert S.shape == S_ref.shape and dX.shape == dX_ref.shape
    assert np.abs(S - S_ref).max() < 1e-4 and np.abs(dX - dX_ref).max() < 1e-4


def test_panel_layout_roundtrip_and_row_blocks():
    from iddgcn_b200.dist import panelize, row_block, unpanelize
    x = torch.arange(2 * 5 * 12, dtype=torch.float32).reshape(2, 5, 12)
    xp = panelize(x, 3)
    assert xp.shape == (3, 2, 5, 4) and xp.is_contiguous()
    assert torch.equal(xp[1, 0, 2], x[0, 2, 4:8])
    assert torch.equal(unpanelize(xp), x)
    assert [row_block(r, 3, 10)[:2] for r in range(3)] == [(0, 4), (4, 8), (8, 10)]
    assert row_block(7, 8, 20_000_000) == (17_500_000, 20_000_000, 2_500_000)
