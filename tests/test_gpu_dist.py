"""Two-GPU NCCL run of the row-partitioned SpMM (skipped on a single-GPU box): concatenated rank outputs must be
bit-identical to the single-GPU kernels on the same graph (same per-row summation order)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_triples

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, R, D, panels, tri, X, dS, out_dir, exchange="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from iddgcn_b200.dist import PartitionedSpMM, row_block
        from iddgcn_b200.graph import Graph
        lo, hi, block = row_block(rank, world, N)
        g = Graph(tri, N, R, row_block=(lo, hi), device=dev)
        rs = dist.new_group(backend="nccl")
        eng = PartitionedSpMM(g, D, world, rank, block, panels=panels, rs_group=rs, exchange=exchange)
        eng.step(torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(dS[:, lo:hi]).to(dev))
        eng.step(None, None)                      # second step reuses resident inputs
        S_local, dX_local = eng.outputs()
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"S_{rank}.npy"), S_local.cpu().numpy())
        np.save(os.path.join(out_dir, f"dX_{rank}.npy"), dX_local.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("N,panels,exchange", [(5001, 4, "nccl"), (4096, 2, "nccl"), (5001, 2, "nvls"), (4096, 4, "nvls")])
def test_partitioned_spmm_two_gpus_matches_single_gpu(tmp_path, N, panels, exchange):
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    R, D, world = 2, 128, 2
    tri = random_triples(N, R, 60000, seed=N, heavy_row=7)
    rng = np.random.default_rng(0)
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    mp.spawn(_worker, args=(world, 29700 + panels + (10 if exchange == "nvls" else 0), N, R, D, panels, tri, X, dS,
                            str(tmp_path), exchange), nprocs=world, join=True)
    g = Graph(tri, N, R)
    S_ref = ops.spmm_fwd(g, torch.from_numpy(X).cuda()).cpu().numpy()
    dX_ref = ops.spmm_bwd_x(g, torch.from_numpy(dS).cuda()).cpu().numpy()
    S = np.concatenate([np.load(tmp_path / f"S_{r}.npy") for r in range(world)], axis=1)
    dX = np.concatenate([np.load(tmp_path / f"dX_{r}.npy") for r in range(world)], axis=0)
    assert np.array_equal(S, S_ref)               # forward: same rows, same order -> bitwise
    # backward sums two rank partials instead of one ordered pass: equal up to fp32 reassociation
    assert np.abs(dX - dX_ref).max() <= 1e-5 * np.abs(dX_ref).max()
