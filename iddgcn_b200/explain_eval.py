"""Explanation ground truth and Precision / Recall / F1 @ top — the host side that follows the explainers.

Mirrors the three reference scripts that sit after GnnExplainer / IDDGCN_explain / explaiNE:

  * `construct_ground_truth`  — explanation/ground_truth.py:6-116 (writes gt_all.csv)
  * `filter_ground_truth`     — explanation/gt_process.py:7-40   (writes gt_filtered_fold{f}.npz, test_filtered_fold{f}.csv)
  * `explanation_metrics`     — explanation/eval_test.py:15-41   (Precision@5, Recall@5, F1@5 per triple)
  * `average_metrics`         — explanation/eval_test.py:52-63

Plain host code (integer tables of a few thousand rows): nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

__all__ = ["construct_ground_truth", "filter_ground_truth", "explanation_metrics", "average_metrics",
           "load_ground_truth", "save_ground_truth"]


def _partner_lists(table: np.ndarray) -> dict:
    """node -> [partners where node is column 0, in file order] + [partners where node is column 2, in file order]."""
    table = np.asarray(table, np.int64).reshape(-1, 3)
    first: dict = {}
    second: dict = {}
    for a, _, b in table.tolist():
        first.setdefault(a, []).append(b)
        second.setdefault(b, []).append(a)
    return {k: first.get(k, []) + second.get(k, []) for k in set(first) | set(second)}


def construct_ground_truth(dc, cc, dd, num_entities: int | None = None) -> list[list[int]]:
    """Ground-truth explanation triples for every drug–mutation triple (ground_truth.py:6-80).

    dc: [n,3] (mutation, rel, drug) triples; cc: [·,3] mutation–mutation similarity pairs; dd: [·,3] drug–drug
    similarity pairs — the rows pandas yields for the three CSV files (the similarity files carry no header line, so the
    reference's `pd.read_csv` drops their first row; pass what `pd.read_csv(...).values` returns).  Returns one row per dc
    triple: [mutation, rel, drug, *flattened (h, r, t) ground-truth triples] in the reference's order.

    Instead of the reference's per-row DataFrame scans, the (drug, mutation, rel) multiplicities are held in one dense
    count table and the three neighbour sweeps are table look-ups."""
    dc = np.asarray(dc, np.int64).reshape(-1, 3)
    n_ent = int(num_entities) if num_entities else int(max(dc[:, [0, 2]].max(), np.asarray(cc)[:, [0, 2]].max(),
                                                          np.asarray(dd)[:, [0, 2]].max())) + 1
    n_rel = int(dc[:, 1].max()) + 1
    count = np.zeros((n_rel, n_ent, n_ent), np.int32)                  # count[rel, drug, mutation]
    np.add.at(count, (dc[:, 1], dc[:, 2], dc[:, 0]), 1)
    mu_partners, drug_partners = _partner_lists(cc), _partner_lists(dd)

    out = []
    for mu, rel, drug in dc.tolist():
        sim_mu = np.asarray(mu_partners.get(mu, []), np.int64)
        sim_drug = np.asarray(drug_partners.get(drug, []), np.int64)
        c = count[rel]
        via_mu = c[drug, sim_mu] if len(sim_mu) else np.zeros(0, np.int32)           # same drug, similar mutation
        via_drug = c[sim_drug, mu] if len(sim_drug) else np.zeros(0, np.int32)       # similar drug, same mutation
        both = (c[np.ix_(sim_drug, sim_mu)] if len(sim_mu) and len(sim_drug)
                else np.zeros((len(sim_drug), len(sim_mu)), np.int32))               # similar drug, similar mutation
        triples = []
        m1 = np.repeat(sim_mu, via_mu)
        triples.append(np.stack([m1, np.full_like(m1, rel), np.full_like(m1, drug)], 1))
        d2 = np.repeat(sim_drug, via_drug)
        triples.append(np.stack([np.full_like(d2, mu), np.full_like(d2, rel), d2], 1))
        dd_idx, mm_idx = np.nonzero(both)                                            # row-major = drug outer, mutation inner
        rep = both[dd_idx, mm_idx]
        d3, m3 = np.repeat(sim_drug[dd_idx], rep), np.repeat(sim_mu[mm_idx], rep)
        triples.append(np.stack([m3, np.full_like(m3, rel), d3], 1))

        # pairs already claimed by the sweeps above (ground_truth.py keeps ONE set for (drug, mutation) and for
        # similarity pairs, so a similarity edge is skipped when its id pair happens to be in it)
        claimed = set()
        if via_mu.any():
            claimed.add((drug, mu))
        claimed.update((int(d), mu) for d in sim_drug[via_drug > 0])
        claimed.update(zip(sim_drug[dd_idx].tolist(), sim_mu[mm_idx].tolist()))
        sim_edges = []
        drug_supported = (via_drug > 0) | (both > 0).any(1)            # similar drug shares the relation somewhere
        for d in sim_drug[drug_supported].tolist():
            if (drug, d) not in claimed:
                sim_edges.append((drug, 2, d))
                claimed.add((drug, d))
        mu_supported = (via_mu > 0) | (both > 0).any(0)
        for m in sim_mu[mu_supported].tolist():
            if (mu, m) not in claimed:
                sim_edges.append((mu, 3, m))
                claimed.add((mu, m))
        flat = np.concatenate(triples + [np.asarray(sim_edges, np.int64).reshape(-1, 3)]).reshape(-1)
        out.append([mu, rel, drug] + flat.tolist())
    return out


def filter_ground_truth(test_triples, gt_rows):
    """gt_process.py:7-40: match the fold's test triples with the ground-truth table on (obj, rel, sbj) and keep those
    with at least one ground-truth triple.  Returns (test_filtered [k,3] int64, ground truth: k lists of [h, r, t] float
    lists — what the reference stores in gt_filtered_fold{f}.npz['result'])."""
    by_key: dict = {}
    for row in gt_rows:
        by_key.setdefault(tuple(int(v) for v in row[:3]), []).append(np.asarray(row[3:], np.float64))
    kept, gts = [], []
    for key in map(tuple, np.asarray(test_triples, np.int64).reshape(-1, 3).tolist()):
        for flat in by_key.get(key, ()):                   # duplicated keys repeat, as in the reference's left merge
            flat = flat[~np.isnan(flat)]
            n = len(flat) // 3
            if n >= 1:
                kept.append(key)
                gts.append(flat[:3 * n].reshape(n, 3).tolist())
    return np.asarray(kept, np.int64).reshape(-1, 3), gts


def explanation_metrics(preds, gt, top: int = 5, exp_num: int = 10):
    """eval_test.py:15-41: (precision, recall, f1) per explained triple.

    preds[i]: [k,3] (h, r, t) explanation triples as the explainers return them; gt[i]: ground-truth triples.  Each
    prediction is looked up as given and reversed (t, r, h).  Two reference behaviours are kept on purpose because the
    published numbers depend on them: the "top" predictions are the first `top` of a Python set of the predictions (hash
    order, not score order), and the true-positive count adds the direct and the reversed hits, so precision can
    exceed 1."""
    results = []
    for pred, truth in zip(preds, gt):
        pred = np.asarray(pred)
        direct = set(map(tuple, pred))
        mirrored = set(map(tuple, pred[:, ::-1]))
        truth = set(map(tuple, truth))
        hits_top = sum(len(set(list(s)[:top]) & truth) for s in (direct, mirrored))
        hits_all = len(direct & truth) + len(mirrored & truth)
        tp, fp, fn = hits_top, top - hits_top, exp_num - hits_all
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0
        f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
        results.append((precision, recall, f1))
    return results


def average_metrics(per_triple, top: int = 5) -> dict:
    """eval_test.py:52-63: column means, keyed like the reference's DataFrame columns."""
    m = np.asarray(per_triple, np.float64).reshape(-1, 3)
    mean = m.mean(0) if len(m) else np.zeros(3)
    return {f"Precision@{top}": float(mean[0]), f"Recall@{top}": float(mean[1]), f"F1@{top}": float(mean[2])}


def save_ground_truth(path, test_filtered, gt) -> None:
    """Ragged, pickle-free container: `triples` [k,3], `flat` [Σ 3·n_i], `offsets` [k+1]."""
    flat = np.asarray([v for g in gt for tri in g for v in tri], np.float64)
    offsets = np.cumsum([0] + [3 * len(g) for g in gt]).astype(np.int64)
    np.savez_compressed(path, triples=np.asarray(test_filtered, np.int64).reshape(-1, 3), flat=flat, offsets=offsets)


def load_ground_truth(path, allow_pickle: bool = False):
    """Reads either the container written by `save_ground_truth` or the reference's gt_filtered_fold{f}.npz (an object
    array under 'result', gt_process.py:39 — NumPy can only open that with allow_pickle=True, which the caller has to
    ask for).  Returns (test triples or None, ground-truth lists)."""
    with np.load(path, allow_pickle=allow_pickle) as z:
        if "offsets" in z.files:
            off, flat = z["offsets"], z["flat"]
            return z["triples"], [flat[off[i]:off[i + 1]].reshape(-1, 3).tolist() for i in range(len(off) - 1)]
        return None, [list(map(list, g)) for g in z["result"]]
