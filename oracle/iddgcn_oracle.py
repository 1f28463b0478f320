"""CPU ORACLE for the IDDGCN hot path — TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (NumPy, fp64 or fp32) of the reference's algorithm for the
relational graph-convolution path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker —
the product (`iddgcn_b200/`) never imports it and has no CPU fallback.

PARITY PINNING.  The reference's arithmetic lives in TensorFlow 2.7.0 (README.md:17-19,
`tensorflow>=2.7.0`; the shipped .h5 files record keras_version 2.7.0), which is not installed and
cannot be installed here, and the reference ships no tests or golden vectors — so this oracle is
"parity unpinned" w.r.t. TensorFlow outputs.  What pins it instead (tests/test_oracle_pins.py):
  * the 20 shipped fold files regenerate bit-exactly through `fold_pipeline` (integer host path);
  * the five shipped trained weight files give the fold metrics of SURVEY.md Appendix D
    (ROC-AUC / AUPR to 4 decimals) through `model_forward` + `eval_protocol`;
  * the hand-written backward (`model_backward`, SURVEY App. A eq. 5-15) agrees with torch autograd
    of the reference's op order (`torch_model_forward`) to ~1e-12 in fp64;
  * the shipped explanation ground-truth table gt_all.csv (1754 ragged rows) regenerates exactly through
    `construct_ground_truth`, and test_filtered_fold4.csv through `filter_ground_truth` (integer host path).
(Running the reference's own pure-Python scripts here — e.g. explanation/eval_test.py, which needs no TensorFlow —
was not possible either: the sandbox does not permit executing code from the reference checkout or unpickling its
object-array .npz files, so `explanation_metrics` is pinned by hand-worked cases only.)

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ------------------------------------------------------------------------------------------------
# adjacency construction — prediction/utils1.py:415-451
# ------------------------------------------------------------------------------------------------


@dataclass
class SparseCOO:
    """What the reference holds as tf.SparseTensor(1, N, N): indices [nnz,3] int64, values [nnz] f32."""
    indices: np.ndarray
    values: np.ndarray
    dense_shape: tuple

    @property
    def nnz(self) -> int:
        return int(self.indices.shape[0])


def distinct(a: np.ndarray) -> np.ndarray:
    """utils1.py:415-417 — np.unique(a, axis=0): lexicographic sort + dedup of rows."""
    return np.unique(a, axis=0)


def get_adj_mats(data, num_entities: int, num_relations: int) -> list[SparseCOO]:
    """utils1.py:420-451.  data: [M,3] (head, rel, tail).  Empty relation -> one explicit 0.0 at (0,0)."""
    data = np.asarray(data)
    out = []
    for i in range(num_relations):
        data_i = data[data[:, 1] == i]                                    # :425
        if not data_i.shape[0]:                                           # :427-429
            indices = np.zeros((1, 2), dtype=np.int64)
            values = np.zeros((1,), dtype=np.float32)
        else:
            indices = distinct(data_i[:, [0, 2]]).astype(np.int64)        # :436-438
            values = np.ones((indices.shape[0],), dtype=np.float32)       # :439
        # tf.sparse.reorder (:447) = row-major order, which np.unique already produced;
        # tf.sparse.reshape to (1,N,N) (:449) prepends a zero batch index.
        order = np.lexsort((indices[:, 1], indices[:, 0]))
        indices = indices[order]
        values = values[order]
        idx3 = np.concatenate([np.zeros((indices.shape[0], 1), np.int64), indices], axis=1)
        out.append(SparseCOO(idx3, values, (1, num_entities, num_entities)))
    return out


def csr_csc_from_adj(adj: list[SparseCOO], num_entities: int, row_lo: int = 0, row_hi: int | None = None):
    """Relation-segmented CSR/CSC in the layout include/iddgcn.h documents (placeholder zeros dropped)."""
    N = num_entities
    row_hi = N if row_hi is None else row_hi
    n_rows = row_hi - row_lo
    R = len(adj)
    rows, cols, rels = [], [], []
    for r, a in enumerate(adj):
        keep = (a.values != 0) & (a.indices[:, 1] >= row_lo) & (a.indices[:, 1] < row_hi)
        rows.append(a.indices[keep, 1] - row_lo)
        cols.append(a.indices[keep, 2])
        rels.append(np.full(int(keep.sum()), r, np.int64))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    rels = np.concatenate(rels) if rels else np.zeros(0, np.int64)
    nnz = rows.shape[0]
    seg = rels * n_rows + rows
    csr_rowptr = np.zeros(R * n_rows + 1, np.int64)
    np.add.at(csr_rowptr, seg + 1, 1)
    csr_rowptr = np.cumsum(csr_rowptr)
    order = np.lexsort((rows, cols, rels))          # sort by (rel, tail, head)
    seg_c = rels[order] * N + cols[order]
    csc_rowptr = np.zeros(R * N + 1, np.int64)
    np.add.at(csc_rowptr, seg_c + 1, 1)
    csc_rowptr = np.cumsum(csc_rowptr)
    return {
        "csr_rowptr": csr_rowptr.astype(np.int32), "csr_col": cols.astype(np.int32), "csr_row": rows.astype(np.int32),
        "csc_rowptr": csc_rowptr.astype(np.int32), "csc_row": rows[order].astype(np.int32),
        "csc_perm": order.astype(np.int32), "nnz": nnz, "rel": rels,
    }


# ------------------------------------------------------------------------------------------------
# sparse kernels — tf.sparse.sparse_dense_matmul (IDDGCN.py:70) and its gradient
# ------------------------------------------------------------------------------------------------


def edge_values(adj: list[SparseCOO], dtype=np.float64) -> list[np.ndarray]:
    return [a.values.astype(dtype) for a in adj]


def spmm_fwd(adj: list[SparseCOO], X: np.ndarray, vals: list[np.ndarray] | None = None) -> np.ndarray:
    """S[r] = A_r · X.  TF-CPU SparseTensorDenseMatMul semantics: out[row] += v * X[col]
    iterating nnz in stored row-major order."""
    N, D = X.shape
    S = np.zeros((len(adj), N, D), dtype=X.dtype)
    for r, a in enumerate(adj):
        v = (a.values if vals is None else vals[r]).astype(X.dtype)
        np.add.at(S[r], a.indices[:, 1], v[:, None] * X[a.indices[:, 2]])
    return S


def spmm_bwd_x(adj: list[SparseCOO], dS: np.ndarray, vals: list[np.ndarray] | None = None) -> np.ndarray:
    """dX = Σ_r A_rᵀ · dS_r (adjoint SpMM; TF sparse_grad)."""
    R, N, D = dS.shape
    dX = np.zeros((N, D), dtype=dS.dtype)
    for r, a in enumerate(adj):
        v = (a.values if vals is None else vals[r]).astype(dS.dtype)
        np.add.at(dX, a.indices[:, 2], v[:, None] * dS[r][a.indices[:, 1]])
    return dX


def spmm_bwd_vals(adj: list[SparseCOO], dS: np.ndarray, X: np.ndarray) -> list[np.ndarray]:
    """∂L/∂values[e] = Σ_d dS[row_e, d] · X[col_e, d] (explaiNE.py:17)."""
    return [np.einsum("ed,ed->e", dS[r][a.indices[:, 1]], X[a.indices[:, 2]]) for r, a in enumerate(adj)]


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------


@dataclass
class LayerParams:
    relation_kernels: np.ndarray   # (R, D, D)   IDDGCN.py:25-30
    self_kernel: np.ndarray        # (D, D)      IDDGCN.py:31-36
    relation_weights: np.ndarray   # (R,)        IDDGCN.py:39-44 — created, never used in call
    W_alpha: np.ndarray            # (D, R)      IDDGCN.py:47-52
    b_alpha: np.ndarray            # (R,)        IDDGCN.py:53-58


@dataclass
class ModelParams:
    emb: np.ndarray                        # (N, D) entity_embeddings   IDDGCN.py:215-225
    layers: list[LayerParams] = field(default_factory=list)
    rel_emb: np.ndarray | None = None      # (R, D) DistMult            IDDGCN.py:92-101

    def flat_list(self) -> list[np.ndarray]:
        """Keras trainable_weights / .h5 load order."""
        out = [self.emb]
        for l in self.layers:
            out += [l.relation_kernels, l.self_kernel, l.relation_weights, l.W_alpha, l.b_alpha]
        out.append(self.rel_emb)
        return out

    @staticmethod
    def from_list(ws: list[np.ndarray], num_layers: int = 3) -> "ModelParams":
        assert len(ws) == 2 + 5 * num_layers, f"expected {2 + 5 * num_layers} arrays, got {len(ws)}"
        layers = [LayerParams(*ws[1 + 5 * i: 6 + 5 * i]) for i in range(num_layers)]
        return ModelParams(ws[0], layers, ws[-1])

    def astype(self, dt) -> "ModelParams":
        return ModelParams.from_list([w.astype(dt) for w in self.flat_list()], len(self.layers))

    def flat_vector(self) -> np.ndarray:
        return np.concatenate([w.reshape(-1) for w in self.flat_list()])


def init_params(N: int, R: int, D: int, num_layers: int = 3, seed: int = 0, scale: float = 1.0,
                dtype=np.float64) -> ModelParams:
    """Initialisers with the reference's distributions (IDDGCN.py:29,35,41,49,55,95,221-224); the RNG
    stream is NumPy's, not TF's.  `scale` multiplies the N(0,1) kernels."""
    g = np.random.default_rng(seed)
    emb = g.uniform(0, 1, (N, D))
    layers = []
    for _ in range(num_layers):
        lim = np.sqrt(6.0 / (D + R))
        layers.append(LayerParams(g.standard_normal((R, D, D)) * scale, g.standard_normal((D, D)) * scale,
                                  g.uniform(-0.05, 0.05, (R,)), g.uniform(-lim, lim, (D, R)), np.zeros(R)))
    return ModelParams(emb, layers, g.standard_normal((R, D)) * scale).astype(dtype)


# ------------------------------------------------------------------------------------------------
# forward — IDDGCN_Layer.call (IDDGCN.py:60-79), DistMult.call (:103-109), wiring (:226-275)
# ------------------------------------------------------------------------------------------------


def sigmoid(x):
    # numerically safe logistic, same values as 1/(1+exp(-x)) where that does not overflow
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def softmax_rows(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def layer_forward(lp: LayerParams, S: np.ndarray, head_idx, head_e, tail_idx, tail_e):
    """IDDGCN.py:60-79 in the reference's op order (gather rows of S_i, then multiply by W_i)."""
    R = S.shape[0]
    head_out = head_e @ lp.self_kernel                                   # :62
    tail_out = tail_e @ lp.self_kernel                                   # :63
    alpha = softmax_rows(head_e @ lp.W_alpha + lp.b_alpha)               # :66
    gate = sigmoid(alpha)                                                # :75 (per relation column)
    for i in range(R):                                                   # :68
        head_update = S[i][head_idx]                                     # :71
        tail_update = S[i][tail_idx]                                     # :72
        head_out = head_out + gate[:, i:i + 1] * (head_update @ lp.relation_kernels[i])   # :76
        tail_out = tail_out + gate[:, i:i + 1] * (tail_update @ lp.relation_kernels[i])   # :77
    return sigmoid(head_out), sigmoid(tail_out), alpha, gate             # :79


def distmult(rel_emb, head_e, rel_idx, tail_e):
    """IDDGCN.py:103-109 (without the leading batch axis)."""
    return sigmoid(np.sum(head_e * rel_emb[rel_idx] * tail_e, axis=-1))


def model_forward(p: ModelParams, adj: list[SparseCOO], head, rel, tail, vals=None, return_cache=False):
    """get_IDDGCN_Model wiring (IDDGCN.py:226-275): every layer aggregates over the raw table `all_e`."""
    E = p.emb
    S = spmm_fwd(adj, E, vals)                         # identical in all layers
    H, T = E[head], E[tail]                            # :226-227
    cache = {"S": S, "H": [H], "T": [T], "alpha": [], "gate": []}
    for lp in p.layers:
        H, T, alpha, gate = layer_forward(lp, S, head, H, tail, T)
        cache["H"].append(H); cache["T"].append(T); cache["alpha"].append(alpha); cache["gate"].append(gate)
    y = distmult(p.rel_emb, H, rel, T)
    return (y, cache) if return_cache else y


# ------------------------------------------------------------------------------------------------
# loss — IDDGCN_Model.train_step (IDDGCN.py:161-168) with Keras BinaryCrossentropy
# ------------------------------------------------------------------------------------------------
BCE_EPS = 1e-7


def bce_mean(y_true, y_pred):
    yp = np.clip(y_pred, BCE_EPS, 1.0 - BCE_EPS)
    return float(np.mean(-(y_true * np.log(yp + BCE_EPS) + (1.0 - y_true) * np.log(1.0 - yp + BCE_EPS))))


def bce_mean_grad(y_true, y_pred):
    """d mean(BCE) / d y_pred; zero where the clip is active (SURVEY App. A eq. 15)."""
    yp = np.clip(y_pred, BCE_EPS, 1.0 - BCE_EPS)
    inside = (y_pred >= BCE_EPS) & (y_pred <= 1.0 - BCE_EPS)
    g = -(y_true / (yp + BCE_EPS) - (1.0 - y_true) / (1.0 - yp + BCE_EPS))
    return np.where(inside, g, 0.0) / y_pred.shape[0]


# ------------------------------------------------------------------------------------------------
# backward — hand-derived (SURVEY App. A eq. 5-15); checked against torch autograd in the tests
# ------------------------------------------------------------------------------------------------


def model_backward(p: ModelParams, adj: list[SparseCOO], head, rel, tail, cache, dy, vals=None,
                   want_dvals=False):
    """Gradients of L w.r.t. every parameter (and optionally the adjacency values) given dL/dy [T]."""
    N, D = p.emb.shape
    R = len(adj)
    S = cache["S"]
    H3, T3 = cache["H"][-1], cache["T"][-1]
    y = distmult(p.rel_emb, H3, rel, T3)
    dq = dy * y * (1.0 - y)                                               # eq. 14
    Rk = p.rel_emb[rel]
    dH = dq[:, None] * Rk * T3
    dT = dq[:, None] * Rk * H3
    d_rel_emb = np.zeros_like(p.rel_emb)
    np.add.at(d_rel_emb, rel, dq[:, None] * H3 * T3)
    dS = np.zeros_like(S)
    layer_grads = []
    for li in range(len(p.layers) - 1, -1, -1):
        lp = p.layers[li]
        Hin, Tin = cache["H"][li], cache["T"][li]
        Hout, Tout = cache["H"][li + 1], cache["T"][li + 1]
        alpha, gate = cache["alpha"][li], cache["gate"][li]
        dPH = dH * Hout * (1.0 - Hout)                                    # eq. 5
        dPT = dT * Tout * (1.0 - Tout)
        d_rk = np.zeros_like(lp.relation_kernels)
        dgate = np.zeros_like(gate)
        for i in range(R):
            ZH = S[i][head] @ lp.relation_kernels[i]
            ZT = S[i][tail] @ lp.relation_kernels[i]
            dgate[:, i] = np.sum(dPH * ZH, axis=1) + np.sum(dPT * ZT, axis=1)          # eq. 6
            gH = gate[:, i:i + 1] * dPH
            gT = gate[:, i:i + 1] * dPT
            dZ = np.zeros((N, D), dtype=S.dtype)                                      # eq. 10
            np.add.at(dZ, head, gH)
            np.add.at(dZ, tail, gT)
            d_rk[i] = S[i].T @ dZ                                                     # eq. 11
            dS[i] += dZ @ lp.relation_kernels[i].T
        dalpha = dgate * gate * (1.0 - gate)
        dz = alpha * (dalpha - np.sum(dalpha * alpha, axis=1, keepdims=True))          # eq. 7
        d_Wa = Hin.T @ dz
        d_ba = dz.sum(axis=0)
        d_ws = Hin.T @ dPH + Tin.T @ dPT                                               # eq. 8
        dH = dPH @ lp.self_kernel.T + dz @ lp.W_alpha.T                                # eq. 9
        dT = dPT @ lp.self_kernel.T
        layer_grads.append(LayerParams(d_rk, d_ws, np.zeros_like(lp.relation_weights), d_Wa, d_ba))
    layer_grads.reverse()
    d_emb = spmm_bwd_x(adj, dS, vals)                                                  # eq. 12
    np.add.at(d_emb, head, dH)
    np.add.at(d_emb, tail, dT)
    grads = ModelParams(d_emb, layer_grads, d_rel_emb)
    if want_dvals:
        return grads, spmm_bwd_vals(adj, dS, p.emb)                                    # eq. 13
    return grads


# ------------------------------------------------------------------------------------------------
# Keras Adam and the train step (IDDGCN.py:123-178)
# ------------------------------------------------------------------------------------------------


def adam_update(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m[:] = b1 * m + (1.0 - b1) * g
    v[:] = b2 * v + (1.0 - b2) * g * g
    theta -= lr_t * m / (np.sqrt(v) + eps)


def train_step(p: ModelParams, adj, pos, neg, state, lr=1e-3, apply_update=True):
    """One IDDGCN_Model.train_step.  pos/neg: [T,3] (h,r,t).  state: dict(t, m, v) over flat_list().
    Returns (bce_before_scaling, grads).  relation_weights gets no gradient and is skipped (App. B.4)."""
    N = p.emb.shape[0]
    yp, cp = model_forward(p, adj, pos[:, 0], pos[:, 1], pos[:, 2], return_cache=True)   # :147-153
    yn, cn = model_forward(p, adj, neg[:, 0], neg[:, 1], neg[:, 2], return_cache=True)   # :154-160
    y_pred = np.concatenate([yp, yn])                                                    # :161
    y_true = np.concatenate([np.ones_like(yp), np.zeros_like(yn)])                       # :162
    loss = bce_mean(y_true, y_pred)                                                      # :166
    dy = bce_mean_grad(y_true, y_pred) * (1.0 / N)                                       # :168
    gp = model_backward(p, adj, pos[:, 0], pos[:, 1], pos[:, 2], cp, dy[: yp.shape[0]])
    gn = model_backward(p, adj, neg[:, 0], neg[:, 1], neg[:, 2], cn, dy[yp.shape[0]:])
    grads = ModelParams.from_list([a + b for a, b in zip(gp.flat_list(), gn.flat_list())], len(p.layers))
    if apply_update:
        state["t"] += 1
        for k, (w, g) in enumerate(zip(p.flat_list(), grads.flat_list())):
            if k >= 1 and k < len(p.flat_list()) - 1 and (k - 1) % 5 == 2:
                continue  # relation_weights: gradient is None in TF -> skipped
            adam_update(w, g, state["m"][k], state["v"][k], state["t"], lr)
    return loss, grads


def new_adam_state(p: ModelParams):
    return {"t": 0, "m": [np.zeros_like(w) for w in p.flat_list()], "v": [np.zeros_like(w) for w in p.flat_list()]}


# ------------------------------------------------------------------------------------------------
# torch autograd restatement in the reference's op order — used to validate model_backward
# ------------------------------------------------------------------------------------------------


def torch_model_forward(tparams: list, adj: list[SparseCOO], head, rel, tail, N: int, vals=None):
    """Same computation with torch CPU ops (sparse COO matmul = tf.sparse.sparse_dense_matmul).
    tparams: flat list of torch tensors in flat_list() order; vals: optional list of per-relation torch
    value tensors (autograd flows into them, explaiNE.py:17)."""
    import torch
    E = tparams[0]
    L = (len(tparams) - 2) // 5
    h = torch.as_tensor(head, dtype=torch.long); t = torch.as_tensor(tail, dtype=torch.long)
    r = torch.as_tensor(rel, dtype=torch.long)
    S = []
    for i, a in enumerate(adj):
        v = torch.as_tensor(a.values, dtype=E.dtype) if vals is None else vals[i]
        A = torch.sparse_coo_tensor(torch.as_tensor(a.indices[:, 1:].T.copy()), v, (N, N))
        S.append(torch.sparse.mm(A, E))
    H, T = E[h], E[t]
    for li in range(L):
        rk, sk, _, Wa, ba = tparams[1 + 5 * li: 6 + 5 * li]
        ho, to = H @ sk, T @ sk
        alpha = torch.softmax(H @ Wa + ba, dim=1)
        for i in range(len(adj)):
            g = torch.sigmoid(alpha[:, i])
            ho = ho + g[:, None] * (S[i][h] @ rk[i])
            to = to + g[:, None] * (S[i][t] @ rk[i])
        H, T = torch.sigmoid(ho), torch.sigmoid(to)
    return torch.sigmoid(torch.sum(H * tparams[-1][r] * T, dim=-1))


# ------------------------------------------------------------------------------------------------
# explainers — explanation/GnnExplainer.py:13-74, IDDGCN_explain.py:16-120, explaiNE.py:12-32
# ------------------------------------------------------------------------------------------------


def get_computation_graph(head, tail, data):
    """GnnExplainer.py:13-24 / IDDGCN_explain.py:16-26: all triples touching head or tail (with duplicates)."""
    data = np.asarray(data)

    def neighbors(n):
        return np.concatenate([data[data[:, 0] == n], data[data[:, 2] == n]], axis=0)

    return np.concatenate([neighbors(head), neighbors(tail)], axis=0)


def structure_loss(counts, target):
    """IDDGCN_explain.py:29-41: mean((target - c_i / Σc)^2)."""
    ratios = counts / counts.sum()
    return float(np.mean((target - ratios) ** 2))


def explainer_loss_and_grad(p: ModelParams, adj, triple, mask_logits, *, structure_target=None):
    """One epoch's loss and d loss / d mask_logits (per-edge logits, list per relation).

    GnnExplainer.py:33-51: masked = adj * sigmoid(mask); loss = -before_pred * log(pred + 1e-5);
    IDDGCN_explain.py:68-73: loss = (pred_loss + structure_loss) / 2.  before_pred does not depend on
    the masks, so it is a constant of the gradient."""
    h = np.array([triple[0]]); r = np.array([triple[1]]); t = np.array([triple[2]])
    base_vals = edge_values(adj, p.emb.dtype)
    sg = [sigmoid(m.astype(p.emb.dtype)) for m in mask_logits]
    vals = [bv * s for bv, s in zip(base_vals, sg)]
    before = model_forward(p, adj, h, r, t)[0]
    pred, cache = model_forward(p, adj, h, r, t, vals=vals, return_cache=True)
    pred_loss = -before * np.log(pred[0] + 1e-5)
    dy = np.array([-before / (pred[0] + 1e-5)])
    loss = pred_loss
    scale = 1.0
    if structure_target is not None:
        counts = np.array([v.sum() for v in vals])
        sl = structure_loss(counts, structure_target)
        loss = (pred_loss + sl) / 2.0
        scale = 0.5
    _, dvals = model_backward(p, adj, h, r, t, cache, dy * scale, vals=vals, want_dvals=True)
    if structure_target is not None:
        tot = counts.sum()
        ratios = counts / tot
        dl_dratio = -2.0 * (structure_target - ratios) / len(counts)
        # d ratio_j / d c_i = (δ_ij - ratio_j) / tot
        dl_dc = (dl_dratio - np.sum(dl_dratio * ratios)) / tot
        dvals = [dv + 0.5 * dl_dc[i] for i, dv in enumerate(dvals)]
    dmask = [dv * bv * s * (1.0 - s) for dv, bv, s in zip(dvals, base_vals, sg)]
    return float(loss), dmask, float(before), float(pred[0])


def explaine_scores(p: ModelParams, adj, triple):
    """explaiNE.py:12-32,84-96: gradient of the prediction w.r.t. every adjacency value."""
    h = np.array([triple[0]]); r = np.array([triple[1]]); t = np.array([triple[2]])
    vals = edge_values(adj, p.emb.dtype)
    pred, cache = model_forward(p, adj, h, r, t, vals=vals, return_cache=True)
    _, dvals = model_backward(p, adj, h, r, t, cache, np.ones(1, p.emb.dtype), vals=vals, want_dvals=True)
    return float(pred[0]), dvals


def top_k_edges(adj, scores, k=10, threshold=None):
    """GnnExplainer.py:54-71 / explaiNE.py:21-32: rank edges by score (descending), return [(h,r,t)], scores."""
    trip, sc = [], []
    for i, (a, s) in enumerate(zip(adj, scores)):
        keep = np.ones(len(s), bool) if threshold is None else (s > threshold)
        idx = a.indices[keep][:, 1:]
        if threshold is not None and idx.sum() == 0:   # GnnExplainer.py:60 skips all-zero index sets
            continue
        trip.append(np.stack([idx[:, 0], np.full(len(idx), i), idx[:, 1]], axis=1))
        sc.append(s[keep])
    if not trip:
        return np.zeros((0, 3), np.int64), np.zeros(0)
    trip = np.concatenate(trip); sc = np.concatenate(sc)
    order = np.argsort(-sc, kind="stable")[:k]
    return trip[order], sc[order]


# ------------------------------------------------------------------------------------------------
# host data pipeline — IDDGCN.py:312-373, utils1.py:616-623,741-754,808-815
# ------------------------------------------------------------------------------------------------


def generate_reverse_triplets(triplets):
    """utils1.py:616-623: (tail, rel, head) for every triple with head != tail."""
    return np.array([(t, r, h) for h, r, t in triplets if h != t])


def fold_pipeline(data_dir, num_folds=5, seed=89):
    """Regenerates what IDDGCN.py:312-373 writes for mode 0: per fold X_train, X_test, neg_X_test (as
    DataFrames exactly as saved) and X_train_neg (array (1,·,3) float64)."""
    import os
    import pandas as pd
    from sklearn.model_selection import KFold
    from sklearn.utils import shuffle
    rp = pd.read_csv(os.path.join(data_dir, "triplets_dc.csv"), header=0)                 # :312
    rp = shuffle(rp, random_state=24)                                                     # :313
    mu = pd.read_csv(os.path.join(data_dir, "mu_similar0.97.csv"), header=0)              # :314 (eats row 1)
    dr = pd.read_csv(os.path.join(data_dir, "drug_similar0.78.csv"), header=0)            # :316
    neg = pd.read_csv(os.path.join(data_dir, "negative_dc_28_1754.csv"), header=0)        # :318
    for df in (rp, mu, dr):
        df.columns = ["obj", "rel", "sbj"]                                                # :321-323
    # utils1.split_pos_triple_into_folds(mode=0), utils1.py:741-754
    dc = rp.sample(frac=1, random_state=seed).reset_index(drop=True)
    cc = mu.sample(frac=1, random_state=seed).reset_index(drop=True)
    dd = dr.sample(frac=1, random_state=seed).reset_index(drop=True)
    cc_dd = pd.concat([cc, dd], axis=0)
    pos_splits = []
    for tr, te in KFold(n_splits=num_folds).split(dc):
        pos_splits.append((pd.concat([dc.iloc[tr], cc_dd], axis=0), dc.iloc[te]))
    neg_splits = [(neg.iloc[tr], neg.iloc[te]) for tr, te in KFold(n_splits=num_folds).split(neg)]   # :808-815
    out = []
    for fold in range(num_folds):                                                         # :333
        X_train_triple, X_test_triple = pos_splits[fold]
        neg_X_train, neg_X_test = neg_splits[fold]
        neg_X_test_filtered = neg_X_test[neg_X_test["rel"].isin([0, 1])]                  # :337
        X_test_triple = X_test_triple.drop(
            X_test_triple[(X_test_triple["rel"] == 2) | (X_test_triple["rel"] == 3)].index)   # :344
        syn_train = pd.DataFrame(generate_reverse_triplets(X_train_triple.to_numpy()))    # :345
        syn_neg_train = pd.DataFrame(generate_reverse_triplets(neg_X_train.to_numpy()))   # :346
        syn_test = pd.DataFrame(generate_reverse_triplets(X_test_triple.to_numpy()))      # :347
        for df in (syn_train, syn_test, syn_neg_train):
            df.columns = ["obj", "rel", "sbj"]
        neg_train_all = pd.concat([neg_X_train, syn_neg_train], axis=0)                   # :355
        X_train = pd.concat([X_train_triple, syn_train], axis=0).astype(np.int64)         # :358
        X_test = pd.concat([X_test_triple, syn_test], axis=0).astype(np.int64)            # :359
        out.append({"X_train": X_train, "X_test": X_test, "neg_X_test": neg_X_test_filtered,
                    "X_train_neg": np.expand_dims(neg_train_all, axis=0)})                # :372
    return out


def get_y_true(X_test_pos, X_test_all):
    """utils1.py:657-663: 1 where the scored triple is among the (de-duplicated) positives."""
    import pandas as pd
    pos = pd.DataFrame(X_test_pos).drop_duplicates()
    merged = pd.merge(pd.DataFrame(X_test_all), pos, indicator=True, how="left")
    return (merged["_merge"] == "both").astype(int).values


def eval_protocol(p: ModelParams, X_train, X_test_pos, X_test_neg, N=845, R=4, threshold=0.5):
    """IDDGCN_eval.py:21-30,49-52,97-119: adjacency from train ∪ test positives, score positives ∪
    negatives, confusion matrix at 0.5, ROC-AUC, AUPR."""
    from sklearn.metrics import auc, confusion_matrix, precision_recall_curve, roc_auc_score
    X_test_neg = np.asarray(X_test_neg).astype(np.int64)
    X_test = np.concatenate([X_test_pos, X_test_neg], axis=0)                             # :28
    adj = get_adj_mats(np.concatenate([X_train, X_test_pos], axis=0), N, R)               # :49-50
    preds = model_forward(p, adj, X_test[:, 0], X_test[:, 1], X_test[:, 2])               # :97-105
    y_true = get_y_true(X_test_pos, X_test)                                               # :109
    y_hat = (preds > threshold).astype(int)
    tn, fp, fn, tp = confusion_matrix(y_true, y_hat).ravel()                              # :111
    prec, reca, _ = precision_recall_curve(y_true, preds)                                 # :118
    return {"tp": int(tp), "tn": int(tn), "fp": int(fp), "fn": int(fn),
            "acc": (tn + tp) / (tn + fp + fn + tp), "roc_auc": float(roc_auc_score(y_true, preds)),
            "aupr": float(auc(reca, prec)), "preds": preds, "y_true": y_true}


def explainer_run(p: ModelParams, adjacency_data, triples, init_value, *, epochs=5, lr=1e-3, threshold=0.2,
                  structure_target=None, shared_optimizer=True, k=10):
    """The per-triple mask-optimisation loop of GnnExplainer.replica_step (GnnExplainer.py:26-74; structure_target=None,
    epochs=5, threshold=0.2) and IDDGCN_explain.wgnnexplainer_step (IDDGCN_explain.py:43-120; structure_target=
    [0.4,0.4,0.1,0.1], epochs=10, threshold=0.15).

    Reference semantics kept: R dense (N,N) mask variables all initialised to `init_value` (GnnExplainer.py:123-134);
    gradients reach a mask only at the sub-graph's edge positions (SparseDenseCwiseMul); ONE Keras Adam shared by all
    triples whose slots and iteration count are never reset while the masks are (GnnExplainer.py:72-73,122) — so with
    shared_optimizer=True the dense Adam state carries over and every entry with stale momentum keeps moving.
    shared_optimizer=False is the batched variant: fresh Adam state per triple (identical for the first triple).
    Returns a list of (top-k triples [k',3] (h,r,t), scores [k'], final loss)."""
    N = p.emb.shape[0]
    R = p.rel_emb.shape[0]
    dt = p.emb.dtype
    init_value = np.asarray(init_value, dt)
    adjacency_data = np.asarray(adjacency_data)
    m = np.zeros((R, N, N), dt); v = np.zeros((R, N, N), dt)
    t_step = 0
    out = []
    for trip in np.asarray(triples):
        if not shared_optimizer:
            m[:] = 0; v[:] = 0; t_step = 0
        masks = np.broadcast_to(init_value, (R, N, N)).copy()           # mask.assign(init_value)
        comp = get_computation_graph(trip[0], trip[2], adjacency_data)  # GnnExplainer.py:27
        adj = get_adj_mats(comp, N, R)                                  # :28
        loss = 0.0
        for _ in range(epochs):                                         # :30
            logits = [masks[i][a.indices[:, 1], a.indices[:, 2]] for i, a in enumerate(adj)]
            loss, dmask, _, _ = explainer_loss_and_grad(p, adj, trip, logits, structure_target=structure_target)
            g = np.zeros((R, N, N), dt)
            for i, a in enumerate(adj):
                np.add.at(g[i], (a.indices[:, 1], a.indices[:, 2]), dmask[i])
            t_step += 1
            for i in range(R):                                          # dense Adam on every mask entry (:52)
                adam_update(masks[i], g[i], m[i], v[i], t_step, lr)
        scores = []
        for i, a in enumerate(adj):
            scores.append(a.values.astype(dt) * sigmoid(masks[i][a.indices[:, 1], a.indices[:, 2]]))   # :56
        tk, sk = top_k_edges(adj, scores, k=k, threshold=threshold)     # :57-71
        out.append((tk, sk, loss))
    return out


# ------------------------------------------------------------------------------------------------
# explanation ground truth and Precision/Recall/F1@5 — explanation/ground_truth.py, gt_process.py, eval_test.py
# ------------------------------------------------------------------------------------------------
def construct_ground_truth(dc, cc, dd):
    """explanation/ground_truth.py:6-80,93-116 — for every drug–mutation triple (mu, rel, drug) the list of supporting
    triples: same-relation neighbours through similar mutations / similar drugs, then the similarity edges themselves
    (rel 2 = drug–drug, rel 3 = mutation–mutation).  dc/cc/dd: int arrays [n,3] holding the rows pandas sees AFTER its
    header handling (the two similarity files have no header line, so `pd.read_csv` swallows their first row,
    ground_truth.py:84-86).  Returns one list per dc row: [mu, rel, drug, *flattened ground-truth triples]."""
    dc = np.asarray(dc, np.int64)
    pair_rels: dict = {}                                   # (drug, mu) -> rels in dc row order
    for mu, rel, drug in dc:
        pair_rels.setdefault((int(drug), int(mu)), []).append(int(rel))

    def partners(table, x):                                # :7-9 / :12-14 — column 2 where column 0 == x, then the mirror
        table = np.asarray(table, np.int64)
        return [int(v) for v in table[table[:, 0] == x, 2]] + [int(v) for v in table[table[:, 2] == x, 0]]

    def hits(drug, mu, rel):                               # rows of dc with this (drug, mu, rel), in file order
        return sum(1 for r in pair_rels.get((drug, mu), ()) if r == rel)

    rows = []
    for mu, rel, drug in dc:
        mu, rel, drug = int(mu), int(rel), int(drug)
        sim_mu, sim_drug = partners(cc, mu), partners(dd, drug)
        gt, added = [], set()
        for m in sim_mu:                                                                  # :25-29
            for _ in range(hits(drug, m, rel)):
                gt.append((m, rel, drug)); added.add((drug, mu))
        for d in sim_drug:                                                                # :32-36
            for _ in range(hits(d, mu, rel)):
                gt.append((mu, rel, d)); added.add((d, mu))
        for d in sim_drug:                                                                # :39-44
            for m in sim_mu:
                for _ in range(hits(d, m, rel)):
                    gt.append((m, rel, d)); added.add((d, m))
        for d in sim_drug:                                                                # :47-59
            if hits(d, mu, rel) and (drug, d) not in added:
                gt.append((drug, 2, d)); added.add((drug, d))
            for m in sim_mu:
                if hits(d, m, rel) and (drug, d) not in added:
                    gt.append((drug, 2, d)); added.add((drug, d))
        for m in sim_mu:                                                                  # :62-76
            if hits(drug, m, rel) and (mu, m) not in added:
                gt.append((mu, 3, m)); added.add((mu, m))
            for d in sim_drug:
                if hits(d, m, rel) and (mu, m) not in added:
                    gt.append((mu, 3, m)); added.add((mu, m))
        rows.append([mu, rel, drug] + [v for tri in gt for v in tri])                     # :93-95
    return rows


def filter_ground_truth(test_triples, gt_rows):
    """explanation/gt_process.py:7-40 — left-merge the fold's test triples with the ground-truth table on
    (obj, rel, sbj) and keep the triples that have at least one ground-truth triple.  Returns (kept test triples
    [k,3], list of k lists of [h, r, t] float lists) — the contents of test_filtered_fold{f}.csv and
    gt_filtered_fold{f}.npz['result']."""
    table: dict = {}
    for row in gt_rows:
        table.setdefault((int(row[0]), int(row[1]), int(row[2])), []).append(row[3:])
    kept, gts = [], []
    for h, r, t in np.asarray(test_triples, np.int64):
        for flat in table.get((int(h), int(r), int(t)), [[]]):          # a left merge repeats duplicated keys
            flat = [float(v) for v in flat if not (isinstance(v, float) and np.isnan(v))]
            if len(flat) // 3 >= 1:                                     # :19-22
                kept.append((int(h), int(r), int(t)))
                gts.append([flat[j:j + 3] for j in range(0, len(flat) - len(flat) % 3, 3)])
    return np.asarray(kept, np.int64).reshape(-1, 3), gts


def explanation_metrics(preds, gt, top=5, exp_num=10):
    """explanation/eval_test.py:15-41 — per-triple (Precision@top, Recall@top, F1@top).  `preds[i]`: [k,3] predicted
    explanation triples (h, r, t); `gt[i]`: the ground-truth triples.  A prediction also counts reversed (t, r, h).
    The reference takes its "top" as the first `top` elements of a Python *set* of the predictions, i.e. in hash
    order, not score order (`list(preds_set)[:top]`, :16-19) — kept as is."""
    out = []
    for pred, g in zip(preds, gt):
        pred = np.asarray(pred)
        flip = np.flip(pred, axis=1)                                                      # :33
        preds_set = set(map(tuple, pred))
        flip_set = set(map(tuple, flip))
        gt_set = set(map(tuple, g))
        tp = len(set(list(preds_set)[:top]) & gt_set) + len(set(list(flip_set)[:top]) & gt_set)
        fp = top - tp
        fn = exp_num - (len(preds_set & gt_set) + len(flip_set & gt_set))
        precision = tp / (tp + fp) if (tp + fp) > 0 else 0
        recall = tp / (tp + fn) if (tp + fn) > 0 else 0
        f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
        out.append((precision, recall, f1))
    return out
