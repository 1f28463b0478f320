#!/usr/bin/env python3
"""Generates tests/golden/* from the read-only reference tree (run in the authoring container only).

    python tests/golden/make_golden.py [/root/reference]

The GPU box has no /root/reference, so everything a test needs from it is frozen here:
  data/                   the four raw triple files the fold pipeline starts from (prediction/IDDGCN.py:312-318) and
                          the explanation test set (explanation/GnnExplainer.py:101) — dataset fixtures, not source
  fold_hashes.json        sha256 of the 20 shipped fold files (known answers for oracle.fold_pipeline)
  weights_fold{f}.npz     the five shipped trained weight sets (Keras .h5 read BY ORDER with iddgcn_b200/h5.py)
  h5_sample_fold4.h5      one original .h5 (parser fixture)
  metrics.json            SURVEY Appendix D fold metrics, recomputed here with the oracle in fp64 and fp32
  preds_fold{1,4}.npz     oracle fp64 predictions / pre-sigmoid logits for the eval triples of folds 1 and 4
  explaiNE_preds_fold4.npz  the shipped explaiNE predictions (288,10,3) + scores (288,10), int16 / float32 re-encoding
                          (explanation/explaiNE.py:98); structural pin only (the weights behind it are not shipped)
  gt_all.npz              the shipped explanation ground-truth table (datasets/explanation_datasets/gt_all.csv, written by
                          explanation/ground_truth.py) as ragged int16 rows: `flat` + `offsets` — the known answer for
                          oracle.construct_ground_truth; only the CSV is read (the shipped *.npz files are pickled
                          object arrays and are not opened)
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from iddgcn_b200 import h5  # noqa: E402
from oracle import iddgcn_oracle as O  # noqa: E402


def main(ref="/root/reference"):
    pd_dir = os.path.join(ref, "datasets", "prediction_datasets")
    ex_dir = os.path.join(ref, "datasets", "explanation_datasets")
    data_out = os.path.join(HERE, "data")
    os.makedirs(data_out, exist_ok=True)
    for fn in ("triplets_dc.csv", "mu_similar0.97.csv", "drug_similar0.78.csv", "negative_dc_28_1754.csv"):
        shutil.copyfile(os.path.join(pd_dir, fn), os.path.join(data_out, fn))
    shutil.copyfile(os.path.join(ex_dir, "test_filtered_fold4.csv"), os.path.join(data_out, "test_filtered_fold4.csv"))
    for fn in os.listdir(data_out):
        os.chmod(os.path.join(data_out, fn), 0o644)

    gt = pd.read_csv(os.path.join(ex_dir, "gt_all.csv")).values                     # float64 with NaN padding
    rows = [r[~np.isnan(r)] for r in gt]
    flat = np.concatenate(rows)
    assert np.array_equal(flat, flat.astype(np.int16)), "ground-truth ids no longer fit int16"
    np.savez_compressed(os.path.join(HERE, "gt_all.npz"), flat=flat.astype(np.int16),
                        offsets=np.cumsum([0] + [len(r) for r in rows]).astype(np.int32))

    # shipped explaiNE output (plain int64 / float32 arrays, no pickle): produced from weights the reference does not
    # ship, so it pins no values — but every prediction must be an edge of the train+test adjacency whose row is the
    # head or the tail of the explained triple (1-hop receptive field), scores positive and descending
    z = np.load(os.path.join(ex_dir, "explaiNE_preds_fold4.npz"), allow_pickle=False)
    assert np.array_equal(z["preds"], z["preds"].astype(np.int16))
    np.savez_compressed(os.path.join(HERE, "explaiNE_preds_fold4.npz"), preds=z["preds"].astype(np.int16),
                        scores=z["scores"].astype(np.float32))

    hashes = {}
    for f in range(5):
        for suffix in ("X_train.csv", "X_test.csv", "neg_X_test.csv", "X_train_neg.npy"):
            fn = f"mode0_fold{f}_{suffix}"
            with open(os.path.join(pd_dir, fn), "rb") as fh:
                hashes[fn] = hashlib.sha256(fh.read()).hexdigest()
    json.dump(hashes, open(os.path.join(HERE, "fold_hashes.json"), "w"), indent=1, sort_keys=True)

    wdir = os.path.join(pd_dir, "weights", "IDDGCN_normal")
    metrics = {}
    for f in range(5):
        path = os.path.join(wdir, f"mode0_fold{f}_epoch5000_learnRate0.001_batchsize100_embdim64_weight.h5")
        ws = h5.read_weight_list(path)
        np.savez_compressed(os.path.join(HERE, f"weights_fold{f}.npz"), **{f"w{k:02d}": w for k, w in enumerate(ws)})
        if f == 4:
            shutil.copyfile(path, os.path.join(HERE, "h5_sample_fold4.h5"))
            os.chmod(os.path.join(HERE, "h5_sample_fold4.h5"), 0o644)
        p = O.ModelParams.from_list(ws)
        Xtr = pd.read_csv(os.path.join(pd_dir, f"mode0_fold{f}_X_train.csv")).values
        Xte = pd.read_csv(os.path.join(pd_dir, f"mode0_fold{f}_X_test.csv")).values
        Xng = pd.read_csv(os.path.join(pd_dir, f"mode0_fold{f}_neg_X_test.csv"), index_col=0).values
        entry = {}
        for name, dt in (("fp64", np.float64), ("fp32", np.float32)):
            m = O.eval_protocol(p.astype(dt), Xtr, Xte, Xng)
            entry[name] = {k: (round(float(v), 6) if isinstance(v, float) else v) for k, v in m.items()
                           if k not in ("preds", "y_true")}
            if name == "fp64" and f in (1, 4):
                np.savez_compressed(os.path.join(HERE, f"preds_fold{f}.npz"), preds=m["preds"], y_true=m["y_true"])
        metrics[str(f)] = entry
        print(f, entry["fp64"])
    json.dump(metrics, open(os.path.join(HERE, "metrics.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(*sys.argv[1:])
