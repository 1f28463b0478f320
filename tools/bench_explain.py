#!/usr/bin/env python3
"""Config 5 timing: edge-mask optimisation (GNNExplainer / IDDGCN_explain) over the 288 shipped fold-4 explanation
triples tiled to --triples, block-diagonal batches of --batch on ONE GPU (lists shard over GPUs with no communication:
run one process per GPU on disjoint slices), beside the oracle's per-triple loop on the host.

    python tools/bench_explain.py [--triples 10000] [--batch 512] [--cpu-triples 2] [--gpus P]

With --gpus P the triple list is cut into P contiguous shards, one process per GPU, no communication; results are
gathered on the host and the wall time of the slowest shard is reported.
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


def shard_worker(gpu, P, args, q):
    import torch
    torch.cuda.set_device(gpu)
    q.put((gpu, run(args, shard=(gpu, P))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triples", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--cpu-triples", type=int, default=2)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    if args.gpus > 1:
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=shard_worker, args=(g, args.gpus, args, q)) for g in range(args.gpus)]
        t0 = time.perf_counter()
        for p in procs:
            p.start()
        res = []
        while len(res) < args.gpus:
            try:
                res.append(q.get(timeout=5))
            except Exception:
                if any(p.exitcode not in (None, 0) for p in procs):
                    for p in procs:
                        p.terminate()
                    raise SystemExit("a shard failed")
        for p in procs:
            p.join()
        res.sort()
        out = {"workload": res[0][1]["workload"].replace("1 GPU", f"{args.gpus} GPUs, contiguous shards"),
               "wall_seconds_incl_startup": time.perf_counter() - t0}
        for name in ("gnnexplainer", "iddgcn_explain"):
            sec = max(r[1][name]["seconds"] for r in res)
            out[name] = {"seconds_slowest_shard": sec, "triples_per_s": args.triples / sec,
                         "per_shard_seconds": [r[1][name]["seconds"] for r in res]}
        print(json.dumps(out))
        return
    print(json.dumps(run(args)))


def run(args, shard=None):
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
    if shard is not None:
        k, P = shard
        q = q[len(q) * k // P: len(q) * (k + 1) // P]
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
        ex.timings = {}
        t0 = time.perf_counter()
        ex.explain(q, init, batch_size=args.batch, **kw)
        out[name]["profiled_run_seconds"] = time.perf_counter() - t0
        out[name]["phase_seconds"] = {k: round(v, 4) for k, v in ex.timings.items()}
        ex.timings = None
    if args.cpu_triples and shard is None:
        from oracle import iddgcn_oracle as O
        p = O.ModelParams.from_list([w.astype(np.float32) for w in ws])
        t0 = time.perf_counter()
        O.explainer_run(p, data, test[: args.cpu_triples], init, epochs=5, threshold=0.2, shared_optimizer=True)
        out["cpu_oracle_gnnexplainer_s_per_triple"] = (time.perf_counter() - t0) / args.cpu_triples
        out["host_cpus"] = os.cpu_count()
    return out


if __name__ == "__main__":
    main()
