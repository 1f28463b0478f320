#!/usr/bin/env python3
"""Layer-level line (optional): one full IDDGCN train step (3 layers, DistMult, BCE, backward, Adam) on the
config-3 graph (1 M nodes, 2 x 10 M edges, D=128) with T = 2^22 positive triples sampled from the edge list plus T/8
random negatives; kernels N(0,1)·0.1/sqrt(D) so that activations do not saturate.

    python tools/bench_layer.py [--triples 4194304] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--per-rel", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--triples", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from iddgcn_b200 import _lib, get_IDDGCN_Model
    from iddgcn_b200.graph import Adjacency, Graph
    N, R, D, T = args.nodes, 2, args.dim, args.triples
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    tri = []
    for r in range(R):
        h = torch.randint(0, N, (args.per_rel,), device=dev, generator=gen)
        t = torch.randint(0, N, (args.per_rel,), device=dev, generator=gen)
        tri.append(torch.stack([h, torch.full_like(h, r), t], 1))
    tri = torch.cat(tri)
    g = Graph(tri, N, R)
    adj = [Adjacency(g, r) for r in range(R)]
    sel = torch.randint(0, tri.shape[0], (T,), device=dev, generator=gen)
    pos = tri[sel]
    Tn = T // 8
    neg = torch.stack([torch.randint(0, N, (Tn,), device=dev, generator=gen),
                       torch.randint(0, R, (Tn,), device=dev, generator=gen),
                       torch.randint(0, N, (Tn,), device=dev, generator=gen)], 1)
    del tri
    model = get_IDDGCN_Model(N, R, D, D, seed=0, init_weights=False)
    rng = np.random.default_rng(0)
    s = 0.1 / np.sqrt(D)
    ws = [rng.random((N, D), dtype=np.float32)]
    for _ in range(3):
        ws += [rng.standard_normal((R, D, D), dtype=np.float32) * s, rng.standard_normal((D, D), dtype=np.float32) * s,
               np.zeros(R, np.float32), rng.standard_normal((D, R), dtype=np.float32) * 0.1, np.zeros(R, np.float32)]
    ws.append(rng.standard_normal((R, D), dtype=np.float32))
    model.set_weights(ws)
    x = [None, pos[:, 0].contiguous(), pos[:, 1].contiguous(), pos[:, 2].contiguous(), adj]
    neg_np = neg.cpu().numpy()
    losses = []
    for _ in range(2):
        losses.append(float(model.train_step((x, None), neg_triples=neg_np)))
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = model.train_step((x, None), neg_triples=neg_np)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    losses.append(float(loss))
    Tt = T + Tn
    flops = 3 * (2 * 2 * Tt * D * D * 3 + 2 * R * N * D * D * 3)      # self-kernel and relation GEMMs, fwd + 2x bwd
    print(json.dumps({"workload": f"train step: N={N}, edges={g.nnz}, R={R}, D={D}, T={T}+{Tn}, 3 layers",
                      "ms_per_step": dt * 1e3, "triples_per_s": Tt / dt, "launches_per_step": (_lib.launch_count() - l0) / args.steps,
                      "approx_dense_TFLOPs_fp32": flops / dt / 1e12, "losses": losses,
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
