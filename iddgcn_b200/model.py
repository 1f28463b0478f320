"""Drop-in for the reference's model assembly and training step.

`get_IDDGCN_Model(num_entities, num_relations, embedding_dim, output_dim, seed, all_feature_matrix, mode, fold)`
keeps the signature of prediction/IDDGCN.py:201 and returns an object used the way the reference scripts use
their Keras model:

    model([ALL_INDICES(1,N), head(1,T), rel(1,T), tail(1,T), ADJ_MATS])   -> (1,T)   IDDGCN.py:147-153
    model.predict(x=[...])                                                 -> ndarray IDDGCN_eval.py:61-69,97-105
    model.load_weights(path.h5) / model.save_weights(path.h5)                         IDDGCN_eval.py:46, IDDGCN.py:195
    model.compile(loss=..., optimizer=Adam(learning_rate=1e-3)); model.fit(x=[...], y=..., epochs=5000, ...)
    model.get_layer('entity_embeddings').get_weights()[0]                             IDDGCN.py:413
    model.trainable_weights                                                           IDDGCN.py:172

All arithmetic runs in libiddgcn_b200 (hand-written sm_100a kernels); torch only owns memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Sequence

import numpy as np
import torch

from . import _lib, h5
from ._lib import VALS_EXPLICIT, VALS_ONES, VALS_SIGMOID, check, load, ptr
from .graph import Adjacency, Graph, collect_edge_values
from .plan import Plan

LAYER_WEIGHT_NAMES = ["relation_kernels:0", "self_kernel:0", "relation_weights:0", "W_alpha:0", "b_alpha:0"]


class _F32DevArray:
    def __init__(self, p: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (p, False), "version": 2}
        self._owner = owner


class Adam:
    """Hyper-parameter holder with the Keras defaults the reference relies on (tf.keras.optimizers.Adam)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon


class History:
    def __init__(self):
        self.history = {"loss": []}


class _LayerView:
    def __init__(self, model, name, idxs):
        self._m, self.name, self._idxs = model, name, idxs

    def get_weights(self):
        w = self._m.get_weights()
        return [w[i] for i in self._idxs]


def _unbatch_idx(x, device):
    """(1,T) / (T,) int array or tensor -> contiguous int64 vector (device tensors stay on device)."""
    if isinstance(x, torch.Tensor):
        return x.reshape(-1).to(torch.int64)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x).reshape(-1)).astype(np.int64, copy=False))


class IDDGCN_Model:
    def __init__(self, num_entities, num_relations, embedding_dim, output_dim, seed, mode=0, fold=0, *,
                 num_layers=3, device=None, neg_weight=1.0, init_weights=True):
        if embedding_dim != output_dim:
            raise ValueError("the reference wiring needs embedding_dim == output_dim (IDDGCN.py:62 on head_e)")
        if not torch.cuda.is_available():
            raise RuntimeError("iddgcn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.num_entities, self.num_relations, self.dim = int(num_entities), int(num_relations), int(embedding_dim)
        self.num_layers, self.seed, self.mode, self.fold, self.neg_weight = num_layers, seed, mode, fold, neg_weight
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = load()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.iddgcn_model_create(self.num_entities, self.num_relations, self.dim, num_layers,
                                                C.byref(handle)))
        self._h = handle
        nw = 2 + 5 * num_layers
        offs, sizes = (C.c_int64 * nw)(), (C.c_int64 * nw)()
        check(self._lib.iddgcn_model_param_layout(self._h, offs, sizes))
        self._offsets, self._sizes = [int(o) for o in offs], [int(s) for s in sizes]
        N, R, D = self.num_entities, self.num_relations, self.dim
        self._shapes = [(N, D)] + [(R, D, D), (D, D), (R,), (D, R), (R,)] * num_layers + [(R, D)]
        total = int(self._lib.iddgcn_model_num_params(self._h))
        ptrs = [C.c_void_p() for _ in range(4)]
        check(self._lib.iddgcn_model_buffers(self._h, *[C.byref(p) for p in ptrs]))
        self.params, self.grads, self.adam_m, self.adam_v = (
            torch.as_tensor(_F32DevArray(p.value, total, self), device=self.device) for p in ptrs)
        self.optimizer = Adam()
        self.step_count = 0
        self._plans: dict = {}
        self.neg_path_template = "../datasets/prediction_datasets/mode{mode}_fold{fold}_X_train_neg.npy"
        if init_weights:          # replicas that get their weights copied in skip the random initialisation
            self._init_weights(seed)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            for p, _keep in list(self._plans.values()):
                p.close()
            self._plans.clear()
            self._lib.iddgcn_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ----------------------------------------------------------------------------------
    def _init_weights(self, seed):
        """Same distributions as the reference initialisers (IDDGCN.py:29,35,41,49,55,95,221-224).  The random
        stream is NumPy's, not TensorFlow's: bit-identical initial weights need `set_weights`/`load_weights`."""
        g = np.random.default_rng(seed)
        N, R, D = self.num_entities, self.num_relations, self.dim
        ws = [g.uniform(0.0, 1.0, (N, D))]
        lim = np.sqrt(6.0 / (D + R))
        for _ in range(self.num_layers):
            ws += [g.standard_normal((R, D, D)), g.standard_normal((D, D)), g.uniform(-0.05, 0.05, (R,)),
                   g.uniform(-lim, lim, (D, R)), np.zeros(R)]
        ws.append(g.standard_normal((R, D)))
        self.set_weights(ws)

    def set_weights(self, weights: Sequence[np.ndarray]):
        if len(weights) != len(self._shapes):
            raise ValueError(f"expected {len(self._shapes)} arrays, got {len(weights)}")
        flat = np.zeros(self.params.numel(), np.float32)
        for w, off, n, shp in zip(weights, self._offsets, self._sizes, self._shapes):
            w = np.asarray(w, np.float32)
            if tuple(w.shape) != tuple(shp):
                raise ValueError(f"weight shape {w.shape} does not match {shp}")
            flat[off:off + n] = w.reshape(-1)
        self.params.copy_(torch.from_numpy(flat))

    def get_weights(self) -> list[np.ndarray]:
        flat = self.params.cpu().numpy()
        return [flat[o:o + n].reshape(s).copy() for o, n, s in zip(self._offsets, self._sizes, self._shapes)]

    def get_gradients(self) -> list[np.ndarray]:
        flat = self.grads.cpu().numpy()
        return [flat[o:o + n].reshape(s).copy() for o, n, s in zip(self._offsets, self._sizes, self._shapes)]

    def weight_tensor(self, k: int) -> torch.Tensor:
        """Zero-copy view of the k-th weight (Keras order) inside the flat parameter buffer."""
        return self.params[self._offsets[k]:self._offsets[k] + self._sizes[k]].view(self._shapes[k])

    @property
    def trainable_weights(self):
        return [self.weight_tensor(k) for k in range(len(self._shapes))]

    def get_layer(self, name):
        if name == "entity_embeddings":
            return _LayerView(self, name, [0])
        if name == "DistMult":
            return _LayerView(self, name, [len(self._shapes) - 1])
        raise KeyError(name)

    def _layer_groups(self):
        ws = self.get_weights()
        groups = [("entity_embeddings", [("entity_embeddings/embeddings:0", ws[0])])]
        for l in range(self.num_layers):
            lname = "rgcn__layer" if l == 0 else f"rgcn__layer_{l}"
            groups.append((lname, list(zip(LAYER_WEIGHT_NAMES, ws[1 + 5 * l:6 + 5 * l]))))
        groups.append(("DistMult", [("DistMult/rel_embedding:0", ws[-1])]))
        return groups

    def load_weights(self, path):
        """Keras HDF5 weight file, loaded BY ORDER like Keras' topological loader (fold 3's layers are named
        rgcn__layer_3..5, so names cannot be trusted)."""
        self.set_weights(h5.read_weight_list(path))
        return self

    def save_weights(self, path):
        h5.write_keras_weights(path, self._layer_groups())

    def reset_states(self):
        pass

    # -- plans / inputs ---------------------------------------------------------------------------
    def _plan(self, head, rel, tail) -> Plan:
        h, r, t = (_unbatch_idx(x, self.device) for x in (head, rel, tail))
        if h.is_cuda:
            # Device index tensors are keyed by address + version, so the cache entry KEEPS the tensors: while it lives
            # the caching allocator cannot hand their addresses to another (same-sized, version-0) tensor, which would
            # otherwise hit the stale plan and silently score the previous call's triples.
            key = ("dev", h.data_ptr(), r.data_ptr(), t.data_ptr(), h.numel(), h._version, r._version, t._version)
            keep = (h, r, t)
        else:
            key = ("host", h.numel(), hash(h.numpy().tobytes()), hash(r.numpy().tobytes()), hash(t.numpy().tobytes()))
            keep = None
        ent = self._plans.get(key)
        if ent is None:
            if len(self._plans) >= 64:
                old = next(iter(self._plans))
                self._plans.pop(old)[0].close()
            p = Plan(h, r, t, self.num_entities, self.num_relations, device=self.device)
            p._models.append(weakref.ref(self))
            ent = (p, keep)
            self._plans[key] = ent
        return ent[0]

    @staticmethod
    def _split_inputs(inputs):
        if len(inputs) == 5 and isinstance(inputs[4], (list, tuple)):
            all_idx, head, rel, tail, adj = inputs
        else:  # flattened form: [all, head, rel, tail, adj_0, ..., adj_{R-1}]
            all_idx, head, rel, tail, *adj = inputs
        return all_idx, head, rel, tail, list(adj)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- forward ----------------------------------------------------------------------------------
    def forward_values(self, graph: Graph, plan: Plan, vals=None, mode=VALS_ONES) -> torch.Tensor:
        y = torch.empty(plan.T, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.iddgcn_model_forward(self._h, graph.handle, plan.handle, ptr(vals), mode, ptr(y),
                                                 self._stream()))
        return y

    def backward_values(self, graph: Graph, plan: Plan, dy: torch.Tensor, vals=None, mode=VALS_ONES, *,
                        zero_grads=True, want_dvals=False):
        dv = torch.empty(graph.nnz, dtype=torch.float32, device=self.device) if want_dvals else None
        dy = dy.to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            check(self._lib.iddgcn_model_backward(self._h, graph.handle, plan.handle, ptr(vals), mode, ptr(dy),
                                                  int(zero_grads), ptr(dv), self._stream()))
        return dv

    def __call__(self, inputs, training=False):
        _all, head, rel, tail, adj = self._split_inputs(inputs)
        graph, vals, mode = collect_edge_values(adj)
        plan = self._plan(head, rel, tail)
        if vals is not None and vals.requires_grad and torch.is_grad_enabled():
            y = _ModelFn.apply(vals, self, graph, plan, mode)
        else:
            y = self.forward_values(graph, plan, vals.detach() if vals is not None else None, mode)
        return y.unsqueeze(0)                                      # tf.expand_dims(score, axis=0), IDDGCN.py:109

    def predict(self, x, **_kw):
        with torch.no_grad():
            return self(x).cpu().numpy()

    # -- training ---------------------------------------------------------------------------------
    def compile(self, loss=None, optimizer=None, **_kw):
        if optimizer is not None:
            lr = getattr(optimizer, "learning_rate", getattr(optimizer, "lr", 1e-3))
            self.optimizer = Adam(float(lr), float(getattr(optimizer, "beta_1", 0.9)),
                                  float(getattr(optimizer, "beta_2", 0.999)), float(getattr(optimizer, "epsilon", 1e-7)))

    def _negatives(self, neg_triples):
        if neg_triples is None:
            path = self.neg_path_template.format(mode=self.mode, fold=self.fold)      # IDDGCN.py:128
            if not os.path.exists(path):
                raise FileNotFoundError(f"negative triples not given and {path} does not exist")
            neg_triples = np.load(path)
        neg = np.asarray(neg_triples)
        if neg.ndim == 3:
            neg = neg[0]
        return neg.astype(np.int64)

    def train_step(self, data, neg_triples=None, apply_update=True):
        """One IDDGCN_Model.train_step (IDDGCN.py:123-178).  data = (x, y) with x as in fit().  Returns the BCE
        (before the 1/N scaling, what the reference prints) as a 0-d CUDA tensor."""
        x = data[0] if isinstance(data, (tuple, list)) and len(data) == 2 and isinstance(data[0], (list, tuple)) else data
        _all, head, rel, tail, adj = self._split_inputs(x)
        graph, vals, mode = collect_edge_values(adj)
        if mode != VALS_ONES:
            raise ValueError("train_step expects the unweighted adjacency (get_adj_mats output)")
        neg = self._negatives(neg_triples)
        pos_plan = self._plan(head, rel, tail)
        neg_plan = self._plan(neg[:, 0], neg[:, 1], neg[:, 2])
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        o = self.optimizer
        step = self.step_count + 1
        with torch.cuda.device(self.device):
            check(self._lib.iddgcn_model_train_step(self._h, graph.handle, pos_plan.handle, neg_plan.handle,
                                                    o.learning_rate, o.beta_1, o.beta_2, o.epsilon, step,
                                                    int(apply_update), ptr(loss), None, self._stream()))
        if apply_update:
            self.step_count = step
        return loss[0]

    def fit(self, x, y=None, epochs=1, batch_size=None, verbose=1, callbacks=(), neg_triples=None, log_every=500,
            **_kw):
        """model.fit of IDDGCN.py:399-412: the dataset has ONE sample (leading dim 1), so an epoch is one train step.
        The step is captured once into a CUDA graph and replayed; callbacks with a `save_epochs` attribute
        (SaveWeightsCallback, IDDGCN.py:181-199) are invoked at exactly those epochs, others at every chunk end."""
        _all, head, rel, tail, adj = self._split_inputs(x)
        graph, vals, mode = collect_edge_values(adj)
        if mode != VALS_ONES:
            raise ValueError("fit expects the unweighted adjacency (get_adj_mats output)")
        neg = self._negatives(neg_triples)
        pos_plan = self._plan(head, rel, tail)
        neg_plan = self._plan(neg[:, 0], neg[:, 1], neg[:, 2])
        hist = History()
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            else:
                cb.model = self
        stops = sorted({e for cb in callbacks for e in getattr(cb, "save_epochs", []) if 0 < e <= epochs} | {epochs})
        o = self.optimizer
        done = 0
        if getattr(self, "_fit_stream", None) is None:   # one stream per model: the tensor-core GEMM keeps a weight-split
            self._fit_stream = torch.cuda.Stream(device=self.device)   # scratch per stream it has seen
        stream = self._fit_stream
        stream.wait_stream(torch.cuda.current_stream(self.device))
        graph._models.append(weakref.ref(self))          # closing the graph drops the step captured on it
        while done < epochs:
            nxt = min([s for s in stops if s > done] + [done + log_every])
            n = nxt - done
            losses = torch.empty(n, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device), torch.cuda.stream(stream):
                check(self._lib.iddgcn_model_train_loop(self._h, graph.handle, pos_plan.handle, neg_plan.handle,
                                                        o.learning_rate, o.beta_1, o.beta_2, o.epsilon,
                                                        self.step_count + 1, n, ptr(losses), stream.cuda_stream))
            stream.synchronize()
            self.step_count += n
            done = nxt
            lv = losses.cpu().numpy().tolist()
            hist.history["loss"] += lv
            if verbose:
                print(f"Epoch {done}/{epochs} - loss: {lv[-1]:.6f}", flush=True)
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(done - 1, {"loss": lv[-1]})
        return hist


class _ModelFn(torch.autograd.Function):
    """y = model(triples | edge values); differentiable w.r.t. the adjacency values (explaiNE.py:17,
    GnnExplainer.py:51).  Parameter gradients are left in `model.grads`."""

    @staticmethod
    def forward(ctx, vals, model, graph, plan, mode):
        v = vals.detach().to(torch.float32).contiguous()
        ctx.model, ctx.graph, ctx.plan, ctx.mode = model, graph, plan, mode
        ctx.save_for_backward(v)
        return model.forward_values(graph, plan, v, mode)

    @staticmethod
    def backward(ctx, dy):
        (v,) = ctx.saved_tensors
        m = ctx.model
        # the activations cached in the handle must be those of this (plan, vals): recompute, then differentiate
        m.forward_values(ctx.graph, ctx.plan, v, ctx.mode)
        dv = m.backward_values(ctx.graph, ctx.plan, dy.contiguous(), v, ctx.mode, zero_grads=True, want_dvals=True)
        return dv, None, None, None, None


def get_IDDGCN_Model(num_entities, num_relations, embedding_dim, output_dim, seed, all_feature_matrix=None, mode=0,
                     fold=0, **kw) -> IDDGCN_Model:
    """Signature of prediction/IDDGCN.py:201.  `all_feature_matrix` is accepted and unused, as in the reference
    (its use is commented out at IDDGCN.py:219)."""
    return IDDGCN_Model(num_entities, num_relations, embedding_dim, output_dim, seed, mode, fold, **kw)
