#!/usr/bin/env python3
"""One-GPU micro-benchmark of the partitioned backward's launches at the per-rank shapes of the 8-GPU 2x4 grid on
config 4 (row block 0 of 2, 32 feature columns): A = peers' node block written locally, B = own node block without
addends / with one addend / accumulating in place.  Isolates the kernel cost of the epilogue addends from the exchange.

    python tools/bench_addends.py [--dim 32] [--blocks 2]
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


def timed(fn, iters=10):
    ts = []
    for i in range(iters + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--blocks", type=int, default=2)
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    N, R, Dc = wl["N"], wl["R"], a.dim
    dev = torch.device("cuda", 0)
    lo, hi, block = row_block(0, a.blocks, N)
    parts = []
    for r in range(R):
        h, t = bench.raw_edges(N, wl["per_rel"], r)
        keep = h < hi
        hh, tt = torch.from_numpy(h[keep]).to(dev), torch.from_numpy(t[keep]).to(dev)
        parts.append(torch.stack([hh, torch.full_like(hh, r), tt], 1))
    g = Graph(torch.cat(parts), N, R, row_block=(lo, hi))
    del parts
    dS = torch.randn((R, hi - lo, Dc), device=dev)
    own = torch.empty((block, Dc), device=dev)
    other = torch.empty((N - block, Dc), device=dev)
    add1 = torch.randn((block, Dc), device=dev)
    add2 = torch.randn((block, Dc), device=dev)
    out = {"shape": {"nodes": N, "rows": hi - lo, "edges": g.nnz, "Dc": Dc},
           "A_peer_block_local_ms": timed(lambda: ops.spmm_bwd_x_rows(g, dS, other, block, N)),
           "B_own_block_no_addend_ms": timed(lambda: ops.spmm_bwd_x_rows(g, dS, own, 0, block)),
           "B_own_block_one_addend_ms": timed(lambda: ops.spmm_bwd_x_rows(g, dS, own, 0, block, addends=[add1])),
           "B_own_block_two_addends_ms": timed(lambda: ops.spmm_bwd_x_rows(g, dS, own, 0, block, addends=[add1, add2])),
           "B_own_block_accumulate_ms": timed(lambda: ops.spmm_bwd_x_rows(g, dS, own, 0, block, addends=[own])),
           "add_minb": os.environ.get("IDDGCN_ADD_MINB", "3")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
