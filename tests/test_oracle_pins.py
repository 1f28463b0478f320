"""CPU suite, part 1: pin the oracle (oracle/iddgcn_oracle.py, oracle/spmm_ref.c) to every known answer the
reference ships for this path: the 20 fold files, and the metrics the five shipped weight files
give under the IDDGCN_eval.py protocol.  Fixtures come from tests/golden (frozen from the
reference tree by tests/golden/make_golden.py)."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

from helpers import random_triples, rel_err
from oracle import iddgcn_oracle as O

# SURVEY.md Appendix D / BASELINE.md §2: (tp, tn, fp, fn, acc, roc_auc, aupr)
APPENDIX_D = {
    0: (619, 284, 67, 83, 0.8575, 0.9072, 0.9406),
    1: (623, 258, 93, 79, 0.8367, 0.8841, 0.9283),
    2: (615, 263, 88, 87, 0.8338, 0.8832, 0.9192),
    3: (623, 278, 73, 79, 0.8557, 0.9148, 0.9491),
    4: (596, 275, 75, 104, 0.8295, 0.9068, 0.9496),
}


@pytest.fixture(scope="module")
def folds(golden_dir):
    return O.fold_pipeline(os.path.join(golden_dir, "data"))


def load_weights(golden_dir, fold):
    z = np.load(os.path.join(golden_dir, f"weights_fold{fold}.npz"))
    return [z[k] for k in sorted(z.files)]


def test_fold_files_regenerate_bit_exactly(folds, golden_dir):
    """prediction/IDDGCN.py:312-373 output == shipped mode0_fold{f}_* files (sha256 of the exact bytes)."""
    hashes = json.load(open(os.path.join(golden_dir, "fold_hashes.json")))
    for f, fd in enumerate(folds):
        for key, kw in (("X_train", dict(index=False)), ("X_test", dict(index=False)),
                        ("neg_X_test", dict(index_label=None))):
            s = io.StringIO()
            fd[key].to_csv(s, **kw)
            assert hashlib.sha256(s.getvalue().encode()).hexdigest() == hashes[f"mode0_fold{f}_{key}.csv"], (f, key)
        b = io.BytesIO()
        np.save(b, fd["X_train_neg"])
        assert hashlib.sha256(b.getvalue()).hexdigest() == hashes[f"mode0_fold{f}_X_train_neg.npy"], f


def test_fold_shapes_match_survey(folds):
    # 37 510 train rows (fold 4: 37 512), per relation at fold 1 [1462,1344,2346,32358]
    assert [len(fd["X_train"]) for fd in folds] == [37510, 37510, 37510, 37510, 37512]
    adj = O.get_adj_mats(folds[1]["X_train"].values, 845, 4)
    assert [a.nnz for a in adj] == [1462, 1344, 2346, 32358]
    # shipped adjacencies are symmetric (reverse triples) — the trap App. C warns about
    for a in adj:
        fwd = set(map(tuple, a.indices[:, 1:]))
        assert fwd == {(t, h) for h, t in fwd}


@pytest.mark.parametrize("fold", [0, 1, 2, 3, 4])
def test_shipped_weights_give_appendix_d_metrics(folds, golden_dir, fold):
    p = O.ModelParams.from_list(load_weights(golden_dir, fold))
    fd = folds[fold]
    stored = json.load(open(os.path.join(golden_dir, "metrics.json")))[str(fold)]
    for name, dt in (("fp64", np.float64), ("fp32", np.float32)):
        m = O.eval_protocol(p.astype(dt), fd["X_train"].values, fd["X_test"].values, fd["neg_X_test"].values)
        tp, tn, fp, fn, acc, auc_, aupr = APPENDIX_D[fold]
        assert (m["tp"], m["tn"], m["fp"], m["fn"]) == (tp, tn, fp, fn)
        assert round(m["acc"], 4) == acc and round(m["roc_auc"], 4) == auc_ and round(m["aupr"], 4) == aupr
        assert abs(m["roc_auc"] - stored[name]["roc_auc"]) < 1e-6
    if fold in (1, 4):
        ref = np.load(os.path.join(golden_dir, f"preds_fold{fold}.npz"))
        m = O.eval_protocol(p.astype(np.float64), fd["X_train"].values, fd["X_test"].values, fd["neg_X_test"].values)
        assert np.array_equal(m["y_true"], ref["y_true"]) and np.allclose(m["preds"], ref["preds"], rtol=0, atol=1e-12)


def test_get_adj_mats_semantics():
    """utils1.py:420-451: dedup, row-major order, (1,N,N) indices, empty relation -> explicit 0.0 at (0,0)."""
    data = np.array([[3, 0, 1], [1, 0, 2], [3, 0, 1], [1, 0, 0], [2, 2, 2], [0, 5, 0]])
    adj = O.get_adj_mats(data, 4, 3)
    assert adj[0].indices.tolist() == [[0, 1, 0], [0, 1, 2], [0, 3, 1]] and adj[0].values.tolist() == [1, 1, 1]
    assert adj[1].indices.tolist() == [[0, 0, 0]] and adj[1].values.tolist() == [0.0]
    assert adj[2].indices.tolist() == [[0, 2, 2]] and adj[0].dense_shape == (1, 4, 4)
    st = O.csr_csc_from_adj(adj, 4)
    assert st["nnz"] == 4 and st["csr_rowptr"].tolist() == [0, 0, 2, 2, 3, 3, 3, 3, 3, 3, 3, 4, 4]
    assert st["csr_col"].tolist() == [0, 2, 1, 2]
    # CSC: (rel, tail, head) order with the CSR edge id of each entry
    assert st["csc_row"].tolist() == [1, 3, 1, 2] and st["csc_perm"].tolist() == [0, 2, 1, 3]


def test_manual_backward_matches_torch_autograd():
    """SURVEY App. A eq. 5-15 (model_backward) vs autograd of the reference op order, fp64, incl. d/d(adj values)."""
    import torch
    rng = np.random.default_rng(0)
    N, D, R, T = 23, 8, 3, 40
    tri = random_triples(N, R, 90, seed=1)
    adj = O.get_adj_mats(tri, N, R)
    p = O.init_params(N, R, D, 3, seed=1, scale=0.3)
    q = np.stack([rng.integers(0, N, T), rng.integers(0, R, T), rng.integers(0, N, T)], 1)
    vals = [rng.uniform(0.2, 1.5, a.nnz) for a in adj]
    y, cache = O.model_forward(p, adj, q[:, 0], q[:, 1], q[:, 2], vals=vals, return_cache=True)
    dy = rng.standard_normal(T)
    g, dv = O.model_backward(p, adj, q[:, 0], q[:, 1], q[:, 2], cache, dy, vals=vals, want_dvals=True)
    tp = [torch.tensor(w, requires_grad=True) for w in p.flat_list()]
    tv = [torch.tensor(v, requires_grad=True) for v in vals]
    ty = O.torch_model_forward(tp, adj, q[:, 0], q[:, 1], q[:, 2], N, vals=tv)
    assert np.abs(ty.detach().numpy() - y).max() < 1e-14
    (ty * torch.tensor(dy)).sum().backward()
    for a, b in zip(g.flat_list(), tp):
        bg = np.zeros_like(a) if b.grad is None else b.grad.numpy()
        assert rel_err(a, bg, scale=np.abs(bg).max() + 1e-30) < 1e-12
    for a, b in zip(dv, tv):
        assert rel_err(a, b.grad.numpy()) < 1e-12


def test_explainer_gradient_matches_finite_differences():
    rng = np.random.default_rng(2)
    N, D, R = 17, 8, 3
    tri = random_triples(N, R, 60, seed=4)
    adj = O.get_adj_mats(tri, N, R)
    p = O.init_params(N, R, D, 3, seed=2, scale=0.3)
    masks = [rng.standard_normal(a.nnz) for a in adj]
    tgt = np.array([0.5, 0.3, 0.2])
    for st in (None, tgt):
        L, dm, before, pred = O.explainer_loss_and_grad(p, adj, tri[0], masks, structure_target=st)
        for r in range(R):
            for e in range(min(2, adj[r].nnz)):
                mp = [m.copy() for m in masks]; mp[r][e] += 1e-6
                mm = [m.copy() for m in masks]; mm[r][e] -= 1e-6
                fd = (O.explainer_loss_and_grad(p, adj, tri[0], mp, structure_target=st)[0] -
                      O.explainer_loss_and_grad(p, adj, tri[0], mm, structure_target=st)[0]) / 2e-6
                assert abs(fd - dm[r][e]) < 1e-6 * max(1.0, abs(fd)) + 1e-9


def test_bce_and_adam_semantics():
    """Keras BinaryCrossentropy clip/eps and Adam update form (App. B.4)."""
    y_true = np.array([1.0, 1.0, 0.0, 0.0])
    y_pred = np.array([1.0, 0.25, 0.0, 0.5])
    e = 1e-7
    expect = -np.mean([np.log(1 - e + e), np.log(0.25 + e), np.log(1 - e + e), np.log(0.5 + e)])
    assert abs(O.bce_mean(y_true, y_pred) - expect) < 1e-12
    g = O.bce_mean_grad(y_true, y_pred)
    assert g[0] == 0 and g[2] == 0 and abs(g[1] + 1 / (0.25 + e) / 4) < 1e-12 and abs(g[3] - 1 / (0.5 + e) / 4) < 1e-12
    th, m, v = np.array([1.0]), np.zeros(1), np.zeros(1)
    O.adam_update(th, np.array([0.5]), m, v, 1)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(th[0] - (1.0 - lr_t * 0.05 / (np.sqrt(0.00025) + 1e-7))) < 1e-15


def test_c_restatement_matches_numpy_oracle():
    from oracle import cref
    N, R, D = 500, 3, 32
    tri = random_triples(N, R, 6000, seed=1, heavy_row=5)
    adj = O.get_adj_mats(tri, N, R)
    st = O.csr_csc_from_adj(adj, N)
    rng = np.random.default_rng(0)
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    S = cref.spmm_csr(st["csr_rowptr"], st["csr_col"], None, X).reshape(R, N, D)
    assert rel_err(S, O.spmm_fwd(adj, X.astype(np.float64))) < 1e-6
    rows = (st["rel"] * N + st["csr_row"]).astype(np.int32)
    S_seq = cref.spmm_coo_seq(rows, st["csr_col"], None, X, R * N).reshape(R, N, D)
    assert np.array_equal(S_seq, S)         # TF-CPU order (one thread, stored order) == per-row order of the OMP port
    dX = cref.spmm_csc_bwd(st["csc_rowptr"], st["csc_row"], st["csc_perm"], None, R, N, N, dS)
    assert rel_err(dX, O.spmm_bwd_x(adj, dS.astype(np.float64))) < 1e-6
    v = rng.uniform(0.5, 1.5, st["nnz"]).astype(np.float32)
    vals = np.split(v.astype(np.float64), np.cumsum([a.nnz for a in adj])[:-1])
    S2 = cref.spmm_csr(st["csr_rowptr"], st["csr_col"], v, X).reshape(R, N, D)
    assert rel_err(S2, O.spmm_fwd(adj, X.astype(np.float64), vals)) < 1e-6
    dX2 = cref.spmm_csc_bwd(st["csc_rowptr"], st["csc_row"], st["csc_perm"], v, R, N, N, dS)
    assert rel_err(dX2, O.spmm_bwd_x(adj, dS.astype(np.float64), vals)) < 1e-6
    # double-accumulator flavour (bench.py's accuracy reference for hub rows): the fp64 oracle rounded once to float32
    S64 = cref.spmm_csr_acc64(st["csr_rowptr"], st["csr_col"], v, X).reshape(R, N, D)
    assert np.array_equal(S64, O.spmm_fwd(adj, X.astype(np.float64), vals).astype(np.float32))


def test_train_step_descends():
    rng = np.random.default_rng(0)
    N, R, D = 30, 3, 8
    tri = random_triples(N, R, 120, seed=3)
    adj = O.get_adj_mats(tri, N, R)
    p = O.init_params(N, R, D, 3, seed=1, scale=0.3)
    neg = np.stack([rng.integers(0, N, 20), rng.integers(0, R, 20), rng.integers(0, N, 20)], 1)
    st = O.new_adam_state(p)
    rw0 = [l.relation_weights.copy() for l in p.layers]
    losses = [O.train_step(p, adj, tri[:50], neg, st, lr=1e-2)[0] for _ in range(6)]
    assert losses[-1] < losses[0]
    assert all(np.array_equal(a, l.relation_weights) for a, l in zip(rw0, p.layers))   # never trained (App. B.4)


def _raw_tables(golden_dir):
    import pandas as pd
    d = os.path.join(golden_dir, "data")
    return tuple(pd.read_csv(os.path.join(d, fn)).values                       # ground_truth.py:84-86 (default header)
                 for fn in ("triplets_dc.csv", "mu_similar0.97.csv", "drug_similar0.78.csv"))


def _shipped_gt_all(golden_dir):
    z = np.load(os.path.join(golden_dir, "gt_all.npz"))
    flat, off = z["flat"], z["offsets"]
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def test_ground_truth_table_regenerates_exactly(golden_dir):
    """explanation/ground_truth.py output == shipped datasets/explanation_datasets/gt_all.csv (all 1754 ragged rows)."""
    rows = O.construct_ground_truth(*_raw_tables(golden_dir))
    assert rows == _shipped_gt_all(golden_dir)


def test_filtered_explanation_set_regenerates_exactly(folds, golden_dir):
    """explanation/gt_process.py on fold 4 == shipped test_filtered_fold4.csv (288 triples, same order)."""
    import pandas as pd
    kept, gts = O.filter_ground_truth(folds[4]["X_test"].values, _shipped_gt_all(golden_dir))
    shipped = pd.read_csv(os.path.join(golden_dir, "data", "test_filtered_fold4.csv")).values
    assert np.array_equal(kept, shipped)
    assert len(gts) == 288 and all(len(g) >= 1 and all(len(t) == 3 for t in g) for g in gts)


def test_explanation_metrics_known_answers():
    """explanation/eval_test.py:15-41 on hand-worked cases (top=5, exp_num=10)."""
    gt = [[[1.0, 0.0, 700.0], [2.0, 0.0, 700.0], [700.0, 2.0, 701.0]]]
    # one direct hit, one reversed hit, one miss: tp=2, fp=3, fn=10-2 -> P=2/5, R=2/10
    pred = [np.array([[1, 0, 700], [701, 2, 700], [5, 1, 6]])]
    (p, r, f), = O.explanation_metrics(pred, gt)
    assert (p, r) == (0.4, 0.2) and f == pytest.approx(2 * 0.4 * 0.2 / 0.6)
    # nothing right: precision 0, recall 0, f1 0 (the guarded divisions)
    assert O.explanation_metrics([np.array([[9, 9, 9]])], gt) == [(0.0, 0.0, 0)]
    # a symmetric triple is counted once directly and once reversed (reference behaviour)
    (p, r, _), = O.explanation_metrics([np.array([[3, 3, 3]])], [[[3.0, 3.0, 3.0]]])
    assert (p, r) == (0.4, 0.2)


def test_shipped_explaine_predictions_structural_pin(folds, golden_dir):
    """explaiNE_preds_fold4.npz (explanation/explaiNE.py:98) comes from weights the reference does not ship, so its
    VALUES pin nothing — its STRUCTURE does: every predicted edge is an entry of the adjacency built from fold-4 train
    + the 288 explained triples (explaiNE.py:59-77: get_adj_mats(concat(train, test))), in the (head, rel, tail) = (row,
    relation, column) convention, and its row is the head or the tail of the explained triple — the model's receptive
    field is one hop, so d(pred)/d(a_e) is zero on every other row.  The oracle's saliency has exactly
    that support, and on the shipped top-10 edges it is non-zero."""
    import pandas as pd
    z = np.load(os.path.join(golden_dir, "explaiNE_preds_fold4.npz"))
    preds, scores = z["preds"].astype(np.int64), z["scores"]
    test = pd.read_csv(os.path.join(golden_dir, "data", "test_filtered_fold4.csv")).values
    assert preds.shape == (288, 10, 3) and scores.shape == (288, 10) and len(test) == 288
    adjacency_data = np.concatenate([folds[4]["X_train"].values, test], 0)
    adj = O.get_adj_mats(adjacency_data, 845, 4)
    edges = {(int(i[1]), r, int(i[2])) for r, a in enumerate(adj) for i in a.indices}
    assert all(tuple(int(x) for x in e) in edges for row in preds for e in row)
    assert np.all(scores > 0) and np.all(np.diff(scores, axis=1) <= 0)
    for i in range(288):
        assert set(preds[i, :, 0]) <= {test[i, 0], test[i, 2]}
    # the oracle's edge-value gradient: support = rows {head, tail}; shipped top-10 edges lie inside it
    p = O.ModelParams.from_list(load_weights(golden_dir, 4))
    for i in (0, 57, 287):
        _, dvals = O.explaine_scores(p, adj, test[i])
        for r, (a, dv) in enumerate(zip(adj, dvals)):
            rows = a.indices[:, 1]
            on = (rows == test[i, 0]) | (rows == test[i, 2])
            assert np.all(dv[~on] == 0)
        shipped = {tuple(int(x) for x in e) for e in preds[i]}
        support = {(int(ix[1]), r, int(ix[2])) for r, a in enumerate(adj) for ix in a.indices
                   if ix[1] in (test[i, 0], test[i, 2])}
        assert shipped <= support
