#!/usr/bin/env python3
"""Explanation quality on fold 4: the three explainers on the GPU, then the reference's Precision / Recall / F1 @5
(explanation/eval_test.py) against the ground truth built from the raw tables (explanation/ground_truth.py →
gt_process.py).  Settings are the reference scripts': GNNExplainer 5 epochs / threshold 0.2 / mask seed 123
(GnnExplainer.py:82-97,123-129), IDDGCN_explain 10 epochs / threshold 0.15 / mask seed 24 / targets [0.4,0.4,0.1,0.1]
(IDDGCN_explain.py:124-141), explaiNE top-10 (explaiNE.py:44).

    python tools/explain_eval.py [--weights file.h5|file.npz] [--data-dir tests/golden/data] [--reference-loop]

--reference-loop runs the mask optimisation one triple at a time with ONE Adam state shared by all triples, which is
what the reference's loop does (the optimizer is created once, GnnExplainer.py:122); the default is the batched engine
with a fresh Adam state per triple.  The mask initialisation is NumPy's, not TensorFlow's `tf.random.normal`, so the
numbers are comparable with the reference's only statistically.
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
    ap.add_argument("--weights", default=os.path.join(ROOT, "tests", "golden", "weights_fold4.npz"))
    ap.add_argument("--data-dir", default=os.path.join(ROOT, "tests", "golden", "data"))
    ap.add_argument("--fold", type=int, default=4)
    ap.add_argument("--batch", type=int, default=288)
    ap.add_argument("--reference-loop", action="store_true")
    ap.add_argument("--out", default=None, help="optional .npz for the predicted explanations")
    args = ap.parse_args()

    import pandas as pd
    import torch
    import train_folds
    from iddgcn_b200 import explain_eval as E
    from iddgcn_b200 import get_IDDGCN_Model
    from iddgcn_b200.explain import Explainer

    raw = [pd.read_csv(os.path.join(args.data_dir, fn)).values
           for fn in ("triplets_dc.csv", "mu_similar0.97.csv", "drug_similar0.78.csv")]
    t0 = time.perf_counter()
    gt_rows = E.construct_ground_truth(*raw)
    fd = train_folds.fold_frames(args.data_dir)[args.fold]
    test, gt = E.filter_ground_truth(np.asarray(fd["X_test"], np.int64), gt_rows)
    t_gt = time.perf_counter() - t0

    m = get_IDDGCN_Model(845, 4, 64, 64, seed=123, fold=args.fold)
    if args.weights.endswith(".npz"):
        z = np.load(args.weights)
        m.set_weights([z[k] for k in sorted(z.files)])
    else:
        m.load_weights(args.weights)
    data = np.concatenate([np.asarray(fd["X_train"], np.int64), test], 0)         # GnnExplainer.py:137
    ex = Explainer(m, data)
    out = {"fold": args.fold, "triples": int(len(test)), "ground_truth_seconds": round(t_gt, 3),
           "mask_optimiser": "one Adam state shared by all triples (reference loop)" if args.reference_loop
           else f"batched, fresh Adam state per triple, batch {args.batch}"}
    saved = {}
    runs = (("GNNExplainer", 123, dict(epochs=5, threshold=0.2)),
            ("IDDGCN_explain", 24, dict(epochs=10, threshold=0.15, structure_target=[0.4, 0.4, 0.1, 0.1])))
    for name, seed, kw in runs:
        init = np.random.default_rng(seed).standard_normal((845, 845)).astype(np.float32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = ex.explain(test, init, batch_size=args.batch, shared_optimizer=args.reference_loop, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        preds = [r[0] for r in res]
        out[name] = dict(E.average_metrics(E.explanation_metrics(preds, gt)), seconds=round(dt, 3),
                         mean_edges_returned=float(np.mean([len(p) for p in preds])))
        saved[name] = preds
    t0 = time.perf_counter()
    res = ex.explaine(test, k=10)
    torch.cuda.synchronize()
    preds = [r[0] for r in res]
    out["explaiNE"] = dict(E.average_metrics(E.explanation_metrics(preds, gt)),
                           seconds=round(time.perf_counter() - t0, 3), mean_edges_returned=10.0)
    saved["explaiNE"] = preds
    if args.out:
        np.savez_compressed(args.out, **{f"{k}_{i}": p for k, v in saved.items() for i, p in enumerate(v)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
