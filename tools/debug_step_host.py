#!/usr/bin/env python3
"""Timeline of GridSpMM.step_host (the e2e leg at N > 1): when each upload / kernel chunk / download finishes,
relative to the step start, from CUDA events on the three streams.  Works at world size 1 (grid 1x1) and under torchrun.

    python tools/debug_step_host.py [--workload c3] [--grid RxC]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from iddgcn_b200.dist import GridSpMM, grid_coords, parse_grid, row_block  # noqa: E402
from iddgcn_b200.graph import Graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--grid", default="auto")
    ap.add_argument("--chunks", type=int, default=4)
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = bench.WORKLOADS[a.workload]
    R, D = wl["R"], wl["D"]
    scale = world if wl["scaling"] == "weak" else 1
    N, per_rel = wl["N"] * scale, wl["per_rel"] * scale
    grid = parse_grid(a.grid, world, D)
    pr, pc = grid_coords(rank, grid)
    lo, hi, _ = row_block(pr, grid[0], N)
    parts = []
    for r in range(R):
        h, t = bench.raw_edges(N, per_rel, r)
        keep = (h >= lo) & (h < hi)
        hh, tt = torch.from_numpy(h[keep]).to(dev), torch.from_numpy(t[keep]).to(dev)
        parts.append(torch.stack([hh, torch.full_like(hh, r), tt], 1))
    g = Graph(torch.cat(parts), N, R, row_block=(lo, hi))
    del parts
    eng = GridSpMM(g, D, world, rank, grid)
    Dc = eng.Dc
    Xh = torch.rand((eng.n_rows, Dc)).pin_memory()
    dSh = torch.randn((R, eng.n_rows, Dc)).pin_memory()
    Sh = torch.empty((R, eng.n_rows, Dc), pin_memory=True)
    dXh = torch.empty((eng.n_rows, Dc), pin_memory=True)
    for _ in range(2):
        eng.step_host(Xh, dSh, Sh, dXh, chunks=a.chunks)
    torch.cuda.synchronize()
    eng.timeline = []
    eng.step_host(Xh, dSh, Sh, dXh, chunks=a.chunks)
    torch.cuda.synchronize()
    t0 = eng.timeline[0][1]
    rows = [(label, t0.elapsed_time(ev)) for label, ev in eng.timeline]
    gb = {"X up": Xh.numel() * 4e-9, "dS up": dSh.numel() * 4e-9}
    print(json.dumps({"rank": rank, "grid": f"{grid[0]}x{grid[1]}", "GB": {"X": Xh.numel() * 4e-9, "dS": dSh.numel() * 4e-9,
                                                                           "S": Sh.numel() * 4e-9, "dX": dXh.numel() * 4e-9},
                      "done_at_ms": {k: round(v, 2) for k, v in rows}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
