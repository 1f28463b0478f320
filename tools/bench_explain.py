#!/usr/bin/env python3
"""Config 5 timing: edge-mask optimisation (GNNExplainer / IDDGCN_explain) over the 288 shipped fold-4 explanation
triples tiled to --triples, block-diagonal batches of --batch on ONE GPU (lists shard over GPUs with no communication:
run one process per GPU on disjoint slices), beside the oracle's per-triple loop on the host.

    python tools/bench_explain.py [--triples 10000] [--batch 512] [--cpu-triples 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triples", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--cpu-triples", type=int, default=2)
    args = ap.parse_args()
    import pandas as pd
    import torch
    import train_folds
    from iddgcn_b200 import get_IDDGCN_Model
    from iddgcn_b200.explain import Explainer
    gd = os.path.join(ROOT, "tests", "golden")
    z = np.load(os.path.join(gd, "weights_fold4.npz"))
    ws = [z[k] for k in sorted(z.files)]
    fd = train_folds.fold_frames(os.path.join(gd, "data"))[4]
    test = pd.read_csv(os.path.join(gd, "data", "test_filtered_fold4.csv")).values.astype(np.int64)
    data = np.concatenate([fd["X_train"], test], 0)
    q = np.tile(test, ((args.triples + len(test) - 1) // len(test), 1))[: args.triples]
    m = get_IDDGCN_Model(845, 4, 64, 64, seed=123)
    m.set_weights(ws)
    ex = Explainer(m, data)
    init = np.random.default_rng(24).standard_normal((845, 845)).astype(np.float32)
    out = {"workload": f"c5: {args.triples} triples (288 shipped fold-4 triples tiled), batch {args.batch}, 1 GPU"}
    for name, kw in (("gnnexplainer", dict(epochs=5, threshold=0.2)),
                     ("iddgcn_explain", dict(epochs=10, threshold=0.15, structure_target=[0.4, 0.4, 0.1, 0.1]))):
        ex.explain(q[: args.batch], init, batch_size=args.batch, **kw)        # warm-up (allocations, replica model)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ex.explain(q, init, batch_size=args.batch, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"seconds": dt, "triples_per_s": len(q) / dt, "mean_edges_returned": float(np.mean([len(s) for _, s, _ in res]))}
    if args.cpu_triples:
        from oracle import iddgcn_oracle as O
        p = O.ModelParams.from_list([w.astype(np.float32) for w in ws])
        t0 = time.perf_counter()
        O.explainer_run(p, data, test[: args.cpu_triples], init, epochs=5, threshold=0.2, shared_optimizer=True)
        out["cpu_oracle_gnnexplainer_s_per_triple"] = (time.perf_counter() - t0) / args.cpu_triples
        out["host_cpus"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
