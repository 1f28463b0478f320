#!/usr/bin/env python3
"""Config 2: five-fold CV training + the IDDGCN_eval.py protocol, folds data-parallel across the visible GPUs.

Mirrors prediction/IDDGCN.py:287-419 (mode 0: regenerate the folds from the raw triples, build the adjacency, train
`epochs` steps with Adam(1e-3)) and prediction/IDDGCN_eval.py:21-122 (adjacency from train ∪ test positives, score
positives ∪ negatives, confusion matrix @0.5, ROC-AUC, AUPR — sklearn on the host, as in the reference).  Folds are
independent: fold f runs on GPU f mod P in its own process; nothing is communicated.

    python tools/train_folds.py --data tests/golden/data --epochs 5000 [--gpus P] [--out weights_dir]

From-scratch training cannot reproduce TensorFlow's initializer streams, so the metrics match the reference's range
(mean ROC-AUC 0.899, AUPR 0.937), not its digits; `--weights DIR` evaluates existing .h5 files.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, R, D = 845, 4, 64


def fold_frames(data_dir):
    """Host data pipeline of IDDGCN.py:312-373 (mode 0).  pandas/sklearn exactly as the reference uses them."""
    import pandas as pd
    from sklearn.model_selection import KFold
    from sklearn.utils import shuffle
    rp = shuffle(pd.read_csv(os.path.join(data_dir, "triplets_dc.csv"), header=0), random_state=24)
    mu = pd.read_csv(os.path.join(data_dir, "mu_similar0.97.csv"), header=0)
    dr = pd.read_csv(os.path.join(data_dir, "drug_similar0.78.csv"), header=0)
    neg = pd.read_csv(os.path.join(data_dir, "negative_dc_28_1754.csv"), header=0)
    for df in (rp, mu, dr):
        df.columns = ["obj", "rel", "sbj"]
    seed = 89
    dc = rp.sample(frac=1, random_state=seed).reset_index(drop=True)
    cc_dd = pd.concat([mu.sample(frac=1, random_state=seed).reset_index(drop=True),
                       dr.sample(frac=1, random_state=seed).reset_index(drop=True)], axis=0)
    rev = lambda a: np.array([(t, r, h) for h, r, t in a if h != t])
    out = []
    for (tr, te), (ntr, nte) in zip(KFold(n_splits=5).split(dc), KFold(n_splits=5).split(neg)):
        x_tr = pd.concat([dc.iloc[tr], cc_dd], axis=0).to_numpy()
        x_te = dc.iloc[te]
        x_te = x_te[~x_te["rel"].isin([2, 3])].to_numpy()
        n_tr, n_te = neg.iloc[ntr].to_numpy(), neg.iloc[nte]
        n_te = n_te[n_te["rel"].isin([0, 1])].to_numpy()
        out.append({"X_train": np.concatenate([x_tr, rev(x_tr)]).astype(np.int64),
                    "X_test": np.concatenate([x_te, rev(x_te)]).astype(np.int64),
                    "neg_train": np.concatenate([n_tr, rev(n_tr)]).astype(np.int64),
                    "neg_test": n_te.astype(np.int64)})
    return out


def evaluate(model, fd):
    import pandas as pd
    from sklearn.metrics import auc, confusion_matrix, precision_recall_curve, roc_auc_score
    from iddgcn_b200 import get_adj_mats
    x_te, n_te = fd["X_test"], fd["neg_test"]
    x_all = np.concatenate([x_te, n_te], 0)
    adj = get_adj_mats(np.concatenate([fd["X_train"], x_te], 0), N, R)                 # IDDGCN_eval.py:49-50
    preds = model.predict([np.arange(N).reshape(1, -1), x_all[None, :, 0], x_all[None, :, 1], x_all[None, :, 2], adj])[0]
    pos = pd.DataFrame(x_te).drop_duplicates()                                         # utils1.get_y_true, :657-663
    y_true = (pd.merge(pd.DataFrame(x_all), pos, indicator=True, how="left")["_merge"] == "both").astype(int).values
    tn, fp, fn, tp = confusion_matrix(y_true, (preds > 0.5).astype(int)).ravel()
    prec, reca, _ = precision_recall_curve(y_true, preds)
    return {"tp": int(tp), "tn": int(tn), "fp": int(fp), "fn": int(fn), "acc": float((tp + tn) / len(y_true)),
            "roc_auc": float(roc_auc_score(y_true, preds)), "aupr": float(auc(reca, prec))}


def run_fold(fold, gpu, args, q):
    import torch
    torch.cuda.set_device(gpu)
    from iddgcn_b200 import Adam, get_adj_mats, get_IDDGCN_Model
    fd = fold_frames(args.data)[fold]
    model = get_IDDGCN_Model(N, R, D, D, seed=89, all_feature_matrix=None, mode=0, fold=fold)
    res = {"fold": fold, "gpu": gpu}
    if args.weights:
        model.load_weights(os.path.join(args.weights, f"mode0_fold{fold}_epoch5000_learnRate0.001_batchsize100_embdim64_weight.h5"))
    else:
        xt = fd["X_train"]
        adj = get_adj_mats(xt, N, R)
        model.compile(loss=None, optimizer=Adam(learning_rate=1e-3))
        t0 = time.perf_counter()
        h = model.fit(x=[np.arange(N).reshape(1, -1), xt[None, :, 0], xt[None, :, 1], xt[None, :, 2], adj],
                      y=np.ones((1, len(xt))), epochs=args.epochs, verbose=0, neg_triples=fd["neg_train"], log_every=10**9)
        torch.cuda.synchronize()
        res.update(train_seconds=time.perf_counter() - t0, final_loss=h.history["loss"][-1], epochs=args.epochs)
        if args.out:
            os.makedirs(args.out, exist_ok=True)
            model.save_weights(os.path.join(args.out, f"mode0_fold{fold}_epoch{args.epochs}_learnRate0.001_batchsize100_embdim64_weight.h5"))
    res.update(evaluate(model, fd))
    q.put(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=os.path.join(ROOT, "tests", "golden", "data"))
    ap.add_argument("--epochs", type=int, default=5000)
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--weights", default=None, help="evaluate .h5 files from this directory instead of training")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.multiprocessing as mp
    P = args.gpus or torch.cuda.device_count()
    assert P > 0, "no CUDA device: there is no CPU fallback"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    t0 = time.perf_counter()
    results = []
    folds = list(range(5))
    while folds:                                   # waves of P folds, one process per GPU
        wave, folds = folds[:P], folds[P:]
        procs = [ctx.Process(target=run_fold, args=(f, i % P, args, q)) for i, f in enumerate(wave)]
        for p in procs:
            p.start()
        got = 0
        while got < len(wave):                     # never block on a child that died
            try:
                results.append(q.get(timeout=5))
                got += 1
            except Exception:
                if any(p.exitcode not in (None, 0) for p in procs):
                    for p in procs:
                        p.terminate()
                    raise SystemExit("a fold process failed")
        for p in procs:
            p.join()
    wall = time.perf_counter() - t0
    results.sort(key=lambda r: r["fold"])
    for r in results:
        print(json.dumps(r))
    print(json.dumps({"folds": 5, "gpus": P, "wall_seconds": wall,
                      "mean_roc_auc": float(np.mean([r["roc_auc"] for r in results])),
                      "sd_roc_auc": float(np.std([r["roc_auc"] for r in results])),
                      "mean_aupr": float(np.mean([r["aupr"] for r in results])),
                      "reference_mean_roc_auc": 0.8992, "reference_mean_aupr": 0.9374}))


if __name__ == "__main__":
    main()
