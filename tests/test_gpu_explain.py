"""GPU parity of the explainer paths (explanation/GnnExplainer.py, IDDGCN_explain.py, explaiNE.py) against the oracle's
restatement of the reference loops: batched block-diagonal execution with fresh Adam state per triple, and the
reference's sequential shared-optimizer semantics."""
import os

import numpy as np
import pytest
import torch

from helpers import random_triples, rel_err
from oracle import iddgcn_oracle as O

pytestmark = pytest.mark.gpu


def _setup(N=40, R=4, D=16, M=400, seed=1, scale=0.3):
    tri = random_triples(N, R, M, seed=seed)
    tri = tri[tri[:, 1] < R]
    p = O.init_params(N, R, D, 3, seed=seed, scale=scale)
    p = O.ModelParams.from_list([w.astype(np.float32).astype(np.float64) for w in p.flat_list()])
    from iddgcn_b200 import get_IDDGCN_Model
    m = get_IDDGCN_Model(N, R, D, D, seed=0)
    m.set_weights(p.flat_list())
    init = np.random.default_rng(seed).standard_normal((N, N)).astype(np.float32)
    return tri, p, m, init


def _compare(got, ref, score_tol=2e-5):
    assert len(got) == len(ref)
    for (gt, gs, gl), (rt, rs, rl) in zip(got, ref):
        assert len(gs) == len(rs)
        assert np.allclose(gs, rs, rtol=0, atol=score_tol)
        assert abs(gl - rl) < 1e-5 * max(1.0, abs(rl))
        # identical ranking wherever the oracle's neighbouring scores are separated by more than the tolerance
        for i in range(len(rs)):
            lo = rs[i] - (rs[i + 1] if i + 1 < len(rs) else -1)
            hi = (rs[i - 1] if i > 0 else 2) - rs[i]
            if min(lo, hi) > 10 * score_tol:
                assert tuple(gt[i]) == tuple(rt[i]), (i, gt[i], rt[i])
        assert sorted(map(tuple, gt)) == sorted(map(tuple, rt)) or np.min(np.diff(np.sort(rs))) < 10 * score_tol


def test_subgraph_extraction_matches_reference_computation_graph():
    from iddgcn_b200.explain import Explainer, _Batch
    tri, p, m, init = _setup()
    ex = Explainer(m, tri)
    q = tri[[0, 5, 17, 17]]
    b = _Batch(ex, q)
    got = b.graph.export()
    N, R = 40, 4
    # oracle: per triple get_computation_graph -> get_adj_mats, shifted block-diagonally
    rows = []
    for k, t in enumerate(q):
        comp = O.get_computation_graph(t[0], t[2], tri)
        for i, a in enumerate(O.get_adj_mats(comp, N, R)):
            keep = a.values != 0
            rows.append(np.stack([a.indices[keep, 1] + k * N, np.full(keep.sum(), i), a.indices[keep, 2] + k * N], 1))
    ref = O.csr_csc_from_adj(O.get_adj_mats(np.concatenate(rows), N * len(q), R), N * len(q))
    assert b.graph.nnz == ref["nnz"]
    for key in ("csr_rowptr", "csr_col", "csc_rowptr", "csc_row", "csc_perm"):
        assert np.array_equal(got[key], ref[key]), key


@pytest.mark.parametrize("structure", [False, True])
@pytest.mark.parametrize("batch_size", [3, 64])
def test_batched_mask_optimisation_matches_oracle(structure, batch_size):
    from iddgcn_b200.explain import Explainer
    tri, p, m, init = _setup()
    ex = Explainer(m, tri)
    q = tri[:7]
    tgt = np.array([0.4, 0.4, 0.1, 0.1]) if structure else None
    kw = dict(epochs=10 if structure else 5, threshold=0.15 if structure else 0.2)
    got = ex.explain(q, init, learning_rate=1e-2, structure_target=tgt, batch_size=batch_size, **kw)
    ref = O.explainer_run(p, tri, q, init.astype(np.float64), lr=1e-2, structure_target=tgt, shared_optimizer=False, **kw)
    _compare(got, ref)


def test_shared_optimizer_semantics_match_reference_loop():
    """Adam slots and step count carry over between triples while masks are reset (GnnExplainer.py:72-73,122)."""
    from iddgcn_b200.explain import Explainer
    tri, p, m, init = _setup(seed=3)
    ex = Explainer(m, tri)
    q = tri[[2, 2, 9, 30]]          # a repeated triple makes the carried-over state visible
    got = ex.explain(q, init, epochs=5, learning_rate=1e-2, threshold=0.2, shared_optimizer=True)
    ref = O.explainer_run(p, tri, q, init.astype(np.float64), epochs=5, lr=1e-2, threshold=0.2, shared_optimizer=True)
    _compare(got, ref)
    fresh = O.explainer_run(p, tri, q, init.astype(np.float64), epochs=5, lr=1e-2, threshold=0.2, shared_optimizer=False)
    assert not np.allclose(ref[1][1], fresh[1][1], atol=1e-6)     # the two semantics really differ


def test_explaine_saliency_matches_oracle():
    from iddgcn_b200.explain import Explainer
    tri, p, m, init = _setup(seed=5)
    ex = Explainer(m, tri)
    q = tri[:5]
    got = ex.explaine(q, k=10)
    adj = O.get_adj_mats(tri, 40, 4)
    for (gt, gs, gp), t in zip(got, q):
        pred, dv = O.explaine_scores(p, adj, t)
        assert rel_err(gp, pred) < 1e-5
        rt, rs = O.top_k_edges(adj, dv, k=10)
        assert np.allclose(gs, rs, rtol=0, atol=1e-5 * np.abs(np.concatenate(dv)).max() + 1e-9)
        nz = np.abs(rs) > 1e-7
        assert [tuple(x) for x in gt[nz]] == [tuple(x) for x in rt[nz]]


def test_fold4_explanation_set_runs_batched(golden_dir):
    """C5 shape: the 288 shipped fold-4 explanation triples on the shipped weights, one block-diagonal batch; the
    first triples agree with the oracle loop (fresh-optimizer variant)."""
    import pandas as pd
    from iddgcn_b200 import get_IDDGCN_Model
    from iddgcn_b200.explain import Explainer
    from test_oracle_pins import load_weights
    ws = load_weights(golden_dir, 4)
    fd = O.fold_pipeline(os.path.join(golden_dir, "data"))[4]
    test = pd.read_csv(os.path.join(golden_dir, "data", "test_filtered_fold4.csv")).values.astype(np.int64)
    assert test.shape == (288, 3)
    data = np.concatenate([fd["X_train"].values, test], 0)              # GnnExplainer.py:136
    m = get_IDDGCN_Model(845, 4, 64, 64, seed=123)
    m.set_weights(ws)
    ex = Explainer(m, data)
    init = np.random.default_rng(24).standard_normal((845, 845)).astype(np.float32)
    got = ex.explain(test, init, epochs=5, learning_rate=1e-3, threshold=0.2, batch_size=288)
    assert len(got) == 288 and all(len(s) <= 10 for _, s, _ in got)
    p = O.ModelParams.from_list([w.astype(np.float64) for w in ws])
    ref = O.explainer_run(p, data, test[:3], init.astype(np.float64), epochs=5, lr=1e-3, threshold=0.2,
                          shared_optimizer=False)
    _compare(got[:3], ref, score_tol=5e-5)


def test_fold4_explanations_through_precision_recall_f1(golden_dir):
    """The stage after the explainers (explanation/eval_test.py): GPU explanations of the shipped fold-4 set scored
    against the ground truth rebuilt from the raw tables.  The product's metric code must agree with the oracle's on
    the GPU predictions, and for the first triples the predictions themselves (hence their metrics) with the oracle
    loop's."""
    import pandas as pd
    from iddgcn_b200 import explain_eval as E
    from iddgcn_b200 import get_IDDGCN_Model
    from iddgcn_b200.explain import Explainer
    from test_oracle_pins import load_weights
    d = os.path.join(golden_dir, "data")
    raw = [pd.read_csv(os.path.join(d, fn)).values for fn in ("triplets_dc.csv", "mu_similar0.97.csv",
                                                               "drug_similar0.78.csv")]
    fd = O.fold_pipeline(d)[4]
    test, gt = E.filter_ground_truth(fd["X_test"].values, E.construct_ground_truth(*raw))
    assert test.shape == (288, 3)
    ws = load_weights(golden_dir, 4)
    data = np.concatenate([fd["X_train"].values, test], 0)
    m = get_IDDGCN_Model(845, 4, 64, 64, seed=123)
    m.set_weights(ws)
    ex = Explainer(m, data)
    init = np.random.default_rng(24).standard_normal((845, 845)).astype(np.float32)
    kw = dict(epochs=10, threshold=0.15, structure_target=[0.4, 0.4, 0.1, 0.1])
    got = ex.explain(test, init, batch_size=288, **kw)
    preds = [g[0] for g in got]
    mine = E.explanation_metrics(preds, gt)
    assert mine == O.explanation_metrics(preds, gt)
    avg = E.average_metrics(mine)
    assert 0.0 < avg["Precision@5"] <= 2.0 and 0.0 < avg["Recall@5"] <= 1.0          # explanations do hit the truth
    p = O.ModelParams.from_list([w.astype(np.float64) for w in ws])
    ref = O.explainer_run(p, data, test[:3], init.astype(np.float64), lr=1e-3, shared_optimizer=False, **kw)
    _compare(got[:3], ref, score_tol=5e-5)
    sne = ex.explaine(test[:4], k=10)
    assert E.explanation_metrics([s[0] for s in sne], gt[:4]) == O.explanation_metrics([s[0] for s in sne], gt[:4])
