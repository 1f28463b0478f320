#!/usr/bin/env python3
"""Config 1/2 timing: IDDGCN train steps/s at the shipped fold shape (N=845, R=4, D=64, T=37 510 + 2 806 negatives),
eager C-ABI steps vs the CUDA-graph loop used by fit(), beside the CPU oracle's step time (fp32 NumPy restatement in
the reference's op order).  HBM fraction is not meaningful at 845 nodes: this line is launch-bound.

    python tools/bench_train.py [--fold 1] [--steps 200] [--cpu-steps 1]
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
    ap.add_argument("--fold", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--cpu-steps", type=int, default=1)
    args = ap.parse_args()
    import torch
    from iddgcn_b200 import _lib, get_adj_mats, get_IDDGCN_Model
    from oracle import iddgcn_oracle as O
    fd = O.fold_pipeline(os.path.join(ROOT, "tests", "golden", "data"))[args.fold]
    Xtr = fd["X_train"].values
    neg = fd["X_train_neg"][0].astype(np.int64)
    N, R, D = 845, 4, 64
    model = get_IDDGCN_Model(N, R, D, D, seed=89, all_feature_matrix=None, mode=0, fold=args.fold)
    adj = get_adj_mats(Xtr, N, R)
    x = [np.arange(N).reshape(1, -1), Xtr[None, :, 0], Xtr[None, :, 1], Xtr[None, :, 2], adj]
    for _ in range(3):
        model.train_step((x, None), neg_triples=neg)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.train_step((x, None), neg_triples=neg)
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / args.steps
    launches = (_lib.launch_count() - l0) / args.steps
    model.fit(x=x, epochs=10, verbose=0, neg_triples=neg)          # captures the graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = model.fit(x=x, epochs=args.steps, verbose=0, neg_triples=neg, log_every=10**9)
    torch.cuda.synchronize()
    graph = (time.perf_counter() - t0) / args.steps
    cpu = None
    if args.cpu_steps:
        p = O.init_params(N, R, D, 3, seed=1, dtype=np.float32)
        adj_o = O.get_adj_mats(Xtr, N, R)
        st = O.new_adam_state(p)
        t0 = time.perf_counter()
        for _ in range(args.cpu_steps):
            O.train_step(p, adj_o, Xtr, neg, st)
        cpu = (time.perf_counter() - t0) / args.cpu_steps
    print(json.dumps({"workload": f"c1 fold {args.fold}: N=845 R=4 D=64 T={len(Xtr)}+{len(neg)}",
                      "eager_ms_per_step": eager * 1e3, "graph_ms_per_step": graph * 1e3,
                      "steps_per_s_graph": 1.0 / graph, "kernel_launches_per_step": launches,
                      "loss_first_last": [h.history["loss"][0], h.history["loss"][-1]],
                      "cpu_oracle_ms_per_step": None if cpu is None else cpu * 1e3, "host_cpus": os.cpu_count(),
                      "epochs5000_seconds_graph": 5000 * graph}))


if __name__ == "__main__":
    main()
