#!/usr/bin/env python3
"""Per-rank kernel times of every grid layout, measured on ONE GPU (no exchange): what a rank of a P_r x P_c grid
computes is the row block 0 of P_r blocks at feature width D / P_c, so its forward / backward kernel times can be
measured without the other ranks.  Used to pick bench.py's default grids (profiles/r02_grid_shapes_*.json).

    python tools/bench_shapes.py --workload c4 --world 8 [--grids 8x1,4x2,2x4,1x8]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from iddgcn_b200 import ops  # noqa: E402
from iddgcn_b200.dist import row_block  # noqa: E402
from iddgcn_b200.graph import Graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--grids", default="")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    R, D = wl["R"], wl["D"]
    scale = a.world if wl["scaling"] == "weak" else 1
    N, per_rel = wl["N"] * scale, wl["per_rel"] * scale
    grids = [tuple(int(x) for x in s.split("x")) for s in a.grids.split(",")] if a.grids else \
        [(p, a.world // p) for p in (1, 2, 4, 8, 16) if a.world % p == 0 and p <= a.world and D % (4 * (a.world // p)) == 0]
    dev = torch.device("cuda", 0)
    raw = [bench.raw_edges(N, per_rel, r) for r in range(R)]
    peak, _ = bench.hbm_peak()
    out = []
    for P_r in sorted({g[0] for g in grids}):
        lo, hi, block = row_block(0, P_r, N)
        parts = []
        for r, (h, t) in enumerate(raw):
            keep = h < hi
            hh, tt = torch.from_numpy(h[keep]).to(dev), torch.from_numpy(t[keep]).to(dev)
            parts.append(torch.stack([hh, torch.full_like(hh, r), tt], 1))
        g = Graph(torch.cat(parts, 0), N, R, row_block=(lo, hi))
        del parts
        torch.cuda.empty_cache()
        for (pr, pc) in [g_ for g_ in grids if g_[0] == P_r]:
            Dc = D // pc
            X = torch.rand((N, Dc), device=dev)
            dS = torch.randn((R, hi - lo, Dc), device=dev)
            S = torch.empty((R, hi - lo, Dc), device=dev)
            dX = torch.empty((N, Dc), device=dev)
            tf, tb = [], []
            for i in range(a.iters + 3):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record(); ops.spmm_fwd(g, X, out=S); e[1].record(); ops.spmm_bwd_x(g, dS, out=dX); e[2].record()
                torch.cuda.synchronize()
                if i >= 3:
                    tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
            f, b = float(np.median(tf)), float(np.median(tb))
            bf, bb = bench.algorithmic_bytes(g.nnz, hi - lo, N, R, Dc)
            exch = (pr - 1) * block * Dc * 4
            rec = {"grid": f"{pr}x{pc}", "rows": hi - lo, "edges_per_rank": g.nnz, "Dc": Dc, "fwd_ms": f, "bwd_ms": b,
                   "kernels_ms": f + b, "fwd_frac": bf / f / 1e6 / peak, "bwd_frac": bb / b / 1e6 / peak,
                   "exchange_bytes_each_way": exch, "exchange_ms_at_770GBps_each_way": exch / 770e6,
                   "ideal_edges_per_s_all_ranks": g.nnz * pr / ((f + b) * 1e-3)}
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del X, dS, S, dX
            torch.cuda.empty_cache()
        g.close()
        del g
    print(json.dumps({"workload": a.workload, "world": a.world, "N": N, "shapes": out}))


if __name__ == "__main__":
    main()
