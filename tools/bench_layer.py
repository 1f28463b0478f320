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
    # ---- per-group roofline: events after every kernel group of one eager step (iddgcn_model_profile) ----------
    import ctypes as C
    lib = _lib.load()
    _lib.check(lib.iddgcn_model_profile(model._h, 1))
    acc = np.zeros(len(GROUPS))
    reps = 3
    for _ in range(reps):
        model.train_step((x, None), neg_triples=neg_np)
        buf = (C.c_float * len(GROUPS))()
        _lib.check(lib.iddgcn_model_profile_read(model._h, buf, len(GROUPS)))
        acc += np.array(list(buf))
    _lib.check(lib.iddgcn_model_profile(model._h, 0))
    acc /= reps
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    model_bf = group_models(N, R, D, Tt, 3, g.nnz, int(model.params.numel()))
    groups = {}
    for i, name in enumerate(GROUPS):
        if name == "start":
            continue
        by, fl = model_bf[name]
        ms = float(acc[i])
        rec = {"ms": ms, "share": ms / float(acc.sum()), "algorithmic_GB": by / 1e9,
               "GBps": by / ms / 1e6 if ms > 0 else None, "hbm_frac": by / ms / 1e6 / peak if ms > 0 else None}
        if fl:
            # 3xTF32: three tensor-core products per fp32 product; dense tf32 peak 1.1 PFLOP/s (B200_PROFILING.md)
            rec.update(fp32_TFLOPs=fl / ms / 1e9, tensor_frac_3xtf32=3 * fl / ms / 1e9 / 1100.0)
        groups[name] = rec
    print(json.dumps({"workload": f"train step: N={N}, edges={g.nnz}, R={R}, D={D}, T={T}+{Tn}, 3 layers",
                      "ms_per_step": dt * 1e3, "triples_per_s": Tt / dt, "launches_per_step": (_lib.launch_count() - l0) / args.steps,
                      "approx_dense_TFLOPs_fp32": flops / dt / 1e12, "losses": losses,
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                      "roofline_by_group": {"peak_GBps": peak, "step_ms_eager_with_events": float(acc.sum()),
                                            "bound": "hbm (every group; GEMM groups also list the tensor-pipe fraction)",
                                            "groups": groups}}))


GROUPS = ["start", "zero", "spmm_fwd", "relation_transform", "gather_embeddings", "self_transform", "combine",
          "distmult_fwd", "bce", "distmult_bwd", "bwd_pre", "dW_self", "dW_alpha", "scatter_dZ", "d_input_transform",
          "bwd_post", "scatter_embedding_grad", "dW_relation", "dS_transform", "spmm_bwd", "adam"]


def group_models(N, R, D, T, L, nnz, n_params):
    """(algorithmic bytes, fp32 flops) per kernel group of one train step; T = positives + negatives.  Gathers count
    one 4D-byte row per access (no cache credit), dense operands once per read / write."""
    row = 4 * D
    TD, ND = T * row, N * row
    return {
        "zero": (4 * n_params + L * R * ND, 0),
        "spmm_fwd": (nnz * (row + 4) + R * ND + R * (N + 1) * 4, 0),
        "relation_transform": (L * R * 2 * ND, L * R * 2 * N * D * D),
        "gather_embeddings": (2 * 2 * TD + 2 * T * 4, 0),
        "self_transform": (L * 2 * 2 * TD, L * 2 * 2 * T * D * D),
        "combine": (L * (TD * (1 + 2 + 2 * R + 2) + T * R * 4 + 2 * T * 4), 0),
        "distmult_fwd": (2 * TD + T * 8, 0),
        "bce": (T * 12, 0),
        "distmult_bwd": (4 * TD + T * 12, 0),
        "bwd_pre": (L * (TD * (2 + 2 + 2 * R + 2) + T * R * 8 + 2 * T * 4), 0),
        "dW_self": (L * 2 * 2 * TD, L * 2 * 2 * T * D * D),
        "dW_alpha": (L * (TD + 2 * T * R * 4), L * 2 * T * D * R),
        "scatter_dZ": (L * (2 * TD + 2 * T * (4 + R * 4) + 2 * R * ND), 0),
        "d_input_transform": (L * 2 * 2 * TD, L * 2 * 2 * T * D * D),
        "bwd_post": (L * (2 * TD + T * R * 4), 0),
        "scatter_embedding_grad": (2 * TD + 2 * T * 4 + 2 * ND, 0),
        "dW_relation": (L * R * 2 * ND, L * R * 2 * N * D * D),
        "dS_transform": (L * R * 2 * ND + (L - 1) * R * ND, L * R * 2 * N * D * D),
        "spmm_bwd": (nnz * (row + 4) + 2 * ND + R * (N + 1) * 4, 0),
        "adam": (28 * n_params, 0),
    }


if __name__ == "__main__":
    main()
