"""N>1 host logic on the CPU: world_size-2 gloo runs of the grid-partitioned SpMM exchange (iddgcn_b200/dist.py).

The CUDA kernels cannot run here, so the rank-local compute is injected from the oracle (tests are allowed to use
it); what is under test is the grid arithmetic (row blocks x column groups, short / empty trailing blocks) and the
all-gather / reduce-scatter plumbing: the reassembled rank outputs must equal the single-process result.
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


def _worker(rank, world, port, N, R, D, grid, tri, X, dS, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iddgcn_b200.dist import GridSpMM, grid_coords, row_block
        pr, pc = grid_coords(rank, grid)
        lo, hi, block = row_block(pr, grid[0], N)
        Dc = D // grid[1]
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

        eng = GridSpMM(g, D, world, rank, grid, fwd_fn=fwd, bwd_fn=bwd, device=torch.device("cpu"))
        cs = slice(pc * Dc, (pc + 1) * Dc)
        eng.step(torch.from_numpy(np.ascontiguousarray(X[lo:hi, cs])),
                 torch.from_numpy(np.ascontiguousarray(dS[:, lo:hi, cs])))
        S_local, dX_local = eng.outputs()
        np.save(os.path.join(out_dir, f"S_{rank}.npy"), S_local.numpy())
        np.save(os.path.join(out_dir, f"dX_{rank}.npy"), dX_local.numpy())
    finally:
        dist.destroy_process_group()


def _assemble(tmp_path, kind, grid, N, lead):
    """Put the rank blocks back together: rows by pr, feature columns by pc."""
    from iddgcn_b200.dist import row_block
    cols = []
    for pc in range(grid[1]):
        rows = [np.load(tmp_path / f"{kind}_{pr * grid[1] + pc}.npy") for pr in range(grid[0])]
        cols.append(np.concatenate(rows, axis=lead))
    return np.concatenate(cols, axis=-1)


@pytest.mark.parametrize("N,grid", [(101, (2, 1)), (64, (1, 2)), (3, (2, 1)), (1, (2, 1))])
def test_grid_spmm_world2_gloo(tmp_path, N, grid):
    R, D, world = 3, 16, 2
    tri = random_triples(N, R, 900, seed=N, heavy_row=2 if N > 2 else 0)
    rng = np.random.default_rng(0)
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    port = 29600 + (os.getpid() % 200) + N % 7 + grid[0]
    mp.spawn(_worker, args=(world, port, N, R, D, grid, tri, X, dS, str(tmp_path)), nprocs=world, join=True)
    adj = O.get_adj_mats(tri, N, R)
    S_ref = O.spmm_fwd(adj, X.astype(np.float64))
    dX_ref = O.spmm_bwd_x(adj, dS.astype(np.float64))
    S = _assemble(tmp_path, "S", grid, N, 1)
    dX = _assemble(tmp_path, "dX", grid, N, 0)
    assert S.shape == S_ref.shape and dX.shape == dX_ref.shape
    assert np.abs(S - S_ref).max() < 1e-4 and np.abs(dX - dX_ref).max() < 1e-4


def test_grid_arithmetic():
    from iddgcn_b200.dist import grid_coords, parse_grid, row_block
    assert [row_block(r, 3, 10)[:2] for r in range(3)] == [(0, 4), (4, 8), (8, 10)]
    assert row_block(7, 8, 20_000_000) == (17_500_000, 20_000_000, 2_500_000)
    assert [row_block(r, 4, 5)[:2] for r in range(4)] == [(0, 2), (2, 4), (4, 5), (5, 5)]   # empty trailing block
    assert parse_grid("auto", 8, 128) == (2, 4) and parse_grid(None, 2, 128) == (1, 2) and parse_grid(None, 4, 128) == (1, 4)
    assert parse_grid("auto", 8, 8) == (4, 2)          # D=8 cannot be cut into 4 column groups of >= 4 floats
    assert parse_grid("8x1", 8, 128) == (8, 1) and parse_grid((2, 2), 4, 64) == (2, 2)
    with pytest.raises(ValueError):
        parse_grid("3x2", 8, 128)
    with pytest.raises(ValueError):
        parse_grid("1x8", 8, 16)
    assert [grid_coords(r, (2, 4)) for r in (0, 3, 4, 7)] == [(0, 0), (0, 3), (1, 0), (1, 3)]
