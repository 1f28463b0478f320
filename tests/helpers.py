"""Shared generators for the parity tests (seeded, asymmetric, with the edge cases SURVEY §4 lists)."""
import numpy as np


def random_triples(N, R, M, seed, *, dup_frac=0.2, heavy_row=None, empty_rel=None, bad_rel=0):
    """[M',3] (head, rel, tail): asymmetric, with duplicate triples, optionally one very heavy row,
    one empty relation, and rows whose relation id is outside [0,R) (ignored by the reference)."""
    g = np.random.default_rng(seed)
    h = g.integers(0, N, M)
    t = g.integers(0, N, M)
    r = g.integers(0, R, M)
    tri = np.stack([h, r, t], 1)
    if heavy_row is not None:
        k = max(M // 4, 1)
        tri[:k, 0] = heavy_row
    nd = int(M * dup_frac)
    if nd:
        tri = np.concatenate([tri, tri[g.integers(0, M, nd)]], 0)
    if empty_rel is not None:
        tri = tri[tri[:, 1] != empty_rel]
    if bad_rel:
        bad = np.stack([g.integers(0, N, bad_rel), np.full(bad_rel, R + 3), g.integers(0, N, bad_rel)], 1)
        tri = np.concatenate([tri, bad], 0)
    g.shuffle(tri, axis=0)
    return tri.astype(np.int64)


def rel_err(a, b, scale=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    s = np.abs(b).max() if scale is None else scale
    return float(np.abs(a - b).max() / (s + 1e-300))
