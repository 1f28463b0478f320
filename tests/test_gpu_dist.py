"""Two-GPU runs of the grid-partitioned SpMM (skipped on a single-GPU box; bench.py's in-run parity_check is the
driver-side evidence there): reassembled rank outputs against the single-GPU kernels on the same graph, for the fused
peer-memory exchange, the NCCL exchange and the pure column split, through the device step and the host-buffer step."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_triples

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, R, D, grid, exchange, tri, X, dS, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from iddgcn_b200.dist import GridSpMM, grid_coords, row_block
        from iddgcn_b200.graph import Graph
        pr, pc = grid_coords(rank, grid)
        lo, hi, block = row_block(pr, grid[0], N)
        Dc = D // grid[1]
        g = Graph(tri, N, R, row_block=(lo, hi), device=dev)
        eng = GridSpMM(g, D, world, rank, grid, exchange=exchange, timeout_ms=5000)
        cs = slice(pc * Dc, (pc + 1) * Dc)
        Xl = torch.from_numpy(np.ascontiguousarray(X[lo:hi, cs])).to(dev)
        dSl = torch.from_numpy(np.ascontiguousarray(dS[:, lo:hi, cs])).to(dev)
        eng.step(Xl, dSl)
        for _ in range(3):                        # later steps reuse resident inputs; epochs advance
            eng.step()
        S_local, dX_local = eng.outputs()
        torch.cuda.synchronize()
        eng.check_timeouts()
        np.save(os.path.join(out_dir, f"S_{rank}.npy"), S_local.cpu().numpy())
        np.save(os.path.join(out_dir, f"dX_{rank}.npy"), dX_local.cpu().numpy())
        # host-buffer step: same numbers through the 3-stream pipeline
        Xh, dSh = Xl.cpu().pin_memory(), dSl.cpu().pin_memory()
        Sh = torch.zeros(eng.S.shape).pin_memory()
        dXh = torch.zeros((eng.n_rows, Dc)).pin_memory()
        eng.S.zero_(); eng.dX.zero_()
        eng.step_host(Xh, dSh, Sh, dXh, chunks=3)
        eng.step_host(Xh, dSh, Sh, dXh, chunks=3, overlap=False)       # downloads held until the uploads have landed
        torch.cuda.synchronize()
        eng.check_timeouts()
        np.save(os.path.join(out_dir, f"Sh_{rank}.npy"), Sh.numpy())
        np.save(os.path.join(out_dir, f"dXh_{rank}.npy"), dXh.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _assemble(tmp_path, kind, grid, lead):
    cols = []
    for pc in range(grid[1]):
        rows = [np.load(tmp_path / f"{kind}_{pr * grid[1] + pc}.npy") for pr in range(grid[0])]
        cols.append(np.concatenate(rows, axis=lead))
    return np.concatenate(cols, axis=-1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("N,grid,exchange", [(5001, (2, 1), "p2p"), (4096, (2, 1), "p2p"), (5001, (2, 1), "nccl"),
                                             (4096, (1, 2), "p2p"), (70001, (2, 1), "p2p")])
def test_grid_spmm_two_gpus_matches_single_gpu(tmp_path, N, grid, exchange):
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    R, D, world = 2, 128, 2
    tri = random_triples(N, R, 12 * N, seed=N, heavy_row=7)
    rng = np.random.default_rng(0)
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    port = 29700 + (N % 50) + (10 if exchange == "nccl" else 0) + grid[1]
    mp.spawn(_worker, args=(world, port, N, R, D, grid, exchange, tri, X, dS, str(tmp_path)), nprocs=world, join=True)
    g = Graph(tri, N, R)
    S_ref = ops.spmm_fwd(g, torch.from_numpy(X).cuda()).cpu().numpy()
    dX_ref = ops.spmm_bwd_x(g, torch.from_numpy(dS).cuda()).cpu().numpy()
    for suffix in ("", "h"):
        S = _assemble(tmp_path, "S" + suffix, grid, 1)
        dX = _assemble(tmp_path, "dX" + suffix, grid, 0)
        # per-row order is the same but the kernel instantiation (lanes per row) may differ with the column width, and
        # the backward sums rank partials instead of one ordered pass: equal up to fp32 reassociation
        assert np.abs(S - S_ref).max() <= 1e-5 * np.abs(S_ref).max()
        assert np.abs(dX - dX_ref).max() <= 1e-5 * np.abs(dX_ref).max()
