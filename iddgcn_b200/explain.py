"""Explainers on the CUDA path: GNNExplainer-style edge-mask optimisation (explanation/GnnExplainer.py:26-74),
its structure-loss variant (explanation/IDDGCN_explain.py:29-120) and explaiNE saliency (explanation/explaiNE.py:12-32).

The reference explains the test triples one at a time.  Their 1-hop computation graphs are independent, so a batch of
B triples is laid out block-diagonally (triple b on node ids [b·N, (b+1)·N), embeddings tiled) and one
forward/backward of the ordinary kernels serves the whole batch; edge masks are per-edge logit vectors whose sigmoid
is fused into the SpMM / SDDMM kernels (never materialised).  Batches are independent, so lists of triples shard over
GPUs with no communication.

Two optimizer semantics (SURVEY §7 "Explainer semantics"):
  * `shared_optimizer=False` (default, batchable): fresh Adam state per triple;
  * `shared_optimizer=True`: the reference's behaviour — ONE Adam whose slots and step count carry over from triple
    to triple while the masks are reset (GnnExplainer.py:72-73,122).  Inherently sequential: triples run one by one
    with dense (R,N,N) Adam slots.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import VALS_ONES, VALS_SIGMOID, check, load, ptr
from .graph import Graph
from .model import IDDGCN_Model
from .plan import Plan


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class _Batch:
    """Block-diagonal computation graphs of B triples + the model replica that evaluates them."""

    def __init__(self, ex: "Explainer", triples: np.ndarray):
        lib, dev = ex._lib, ex.device
        self.B = B = len(triples)
        N, R = ex.N, ex.R
        self.triples = triples
        h = torch.from_numpy(np.ascontiguousarray(triples[:, 0])).to(dev)
        t = torch.from_numpy(np.ascontiguousarray(triples[:, 2])).to(dev)
        r = torch.from_numpy(np.ascontiguousarray(triples[:, 1])).to(dev)
        M = C.c_int64()
        st = _stream(dev)
        check(lib.iddgcn_subgraph_triples(ex.graph.handle, ptr(h), ptr(t), B, None, 0, C.byref(M), st))
        tri = torch.empty((max(M.value, 1), 3), dtype=torch.int64, device=dev)
        check(lib.iddgcn_subgraph_triples(ex.graph.handle, ptr(h), ptr(t), B, ptr(tri), M.value, C.byref(M), st))
        self.graph = Graph(tri[:M.value], N * B, R, device=dev)
        shift = torch.arange(B, device=dev, dtype=torch.int64) * N
        self.plan = Plan(h + shift, r, t + shift, N * B, R, device=dev)
        self.model = ex._replica(B)
        import weakref
        self.plan._models.append(weakref.ref(self.model))   # closing the plan drops the activations cached for it
        d = self.graph.device_arrays()
        rows = d["csr_row"].to(torch.int64)
        cols = d["csr_col"].to(torch.int64)
        self.edge_block = rows // N
        self.edge_head = rows % N
        self.edge_tail = cols % N
        offs = torch.from_numpy(self.graph.rel_offsets).to(dev)
        self.edge_rel = torch.bucketize(torch.arange(self.graph.nnz, device=dev), offs[1:], right=True)

    def close(self):
        """Release the device handles now instead of whenever the garbage collector gets to them."""
        self.plan.close()
        self.graph._dev_cache = {}
        self.graph.close()


class Explainer:
    def __init__(self, model: IDDGCN_Model, adjacency_data, *, device=None):
        """adjacency_data: [M,3] triples the computation graphs are drawn from (train ∪ test, GnnExplainer.py:136)."""
        self.src = model
        self.N, self.R, self.D = model.num_entities, model.num_relations, model.dim
        self.device = model.device if device is None else torch.device(device)
        self._lib = load()
        self.graph = Graph(adjacency_data, self.N, self.R, device=self.device)
        self._replicas: dict[int, IDDGCN_Model] = {}
        self.timings = None      # set to {} to collect a per-phase wall-clock breakdown (adds synchronisations)

    def _replica(self, B: int) -> IDDGCN_Model:
        """Model over B·N entities whose embedding table is the trained one tiled B times."""
        m = self._replicas.get(B)
        if m is None:
            if len(self._replicas) >= 2:
                self._replicas.pop(next(iter(self._replicas))).close()
            m = IDDGCN_Model(self.N * B, self.R, self.D, self.D, seed=0, num_layers=self.src.num_layers,
                             device=self.device, init_weights=False)
            self._replicas[B] = m
        src = self.src
        m.weight_tensor(0).view(B, self.N, self.D).copy_(src.weight_tensor(0).unsqueeze(0).expand(B, -1, -1))
        for k in range(1, len(src._shapes)):
            m.weight_tensor(k).copy_(src.weight_tensor(k))
        return m

    # -- mask optimisation --------------------------------------------------------------------------
    def _optimise(self, batch: _Batch, logits, epochs, lr, targets, adam_state):
        lib, dev = self._lib, self.device
        g, plan, model = batch.graph, batch.plan, batch.model
        B, N, R = batch.B, self.N, self.R
        st = _stream(dev)
        before = model.forward_values(g, plan)                              # model([... adj_mats]), GnnExplainer.py:34-40
        tg = torch.tensor(np.asarray(targets, np.float32), device=dev) if targets is not None else None
        c = torch.empty((R, B), dtype=torch.float32, device=dev)
        coef = torch.empty((R, B), dtype=torch.float32, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        dy = torch.empty(B, dtype=torch.float32, device=dev)
        for _ in range(epochs):
            pred = model.forward_values(g, plan, logits, VALS_SIGMOID)      # model([... masked_adjs]), :41-47
            if tg is not None:
                check(lib.iddgcn_block_relation_sums(g.handle, ptr(logits), VALS_SIGMOID, N, B, ptr(c), st))
            check(lib.iddgcn_explain_loss(ptr(before), ptr(pred), ptr(c) if tg is not None else None, ptr(tg), R, B,
                                          ptr(loss), ptr(dy), ptr(coef), st))
            dmask = model.backward_values(g, plan, dy, logits, VALS_SIGMOID, want_dvals=True)   # tape.gradient, :51
            if tg is not None:
                check(lib.iddgcn_structure_grad(g.handle, ptr(logits), VALS_SIGMOID, ptr(coef), N, B, ptr(dmask), st))
            adam_state.apply(logits, dmask, batch, lr)                      # optimizer.apply_gradients, :52
        return before, loss

    def _topk(self, batch: _Batch, logits, threshold, k):
        """threshold σ(mask)·adj, drop relations whose kept indices sum to 0, top-k by score (GnnExplainer.py:54-71)."""
        score = torch.sigmoid(logits)
        keep = score > threshold
        B = batch.B
        blk, rel = batch.edge_block, batch.edge_rel
        idx_sum = torch.zeros((B, self.R), dtype=torch.int64, device=self.device)
        idx_sum.index_put_((blk[keep], rel[keep]), (batch.edge_head + batch.edge_tail)[keep], accumulate=True)
        keep &= idx_sum[blk, rel] != 0
        # order inside a triple: by relation then CSR order — the reference's concat order; stable sort on -score
        order = torch.argsort(torch.where(keep, -score, torch.full_like(score, float("inf"))), stable=True)
        order = order[torch.argsort(blk[order], stable=True)]
        blk_s = blk[order].cpu().numpy()
        keep_s = keep[order].cpu().numpy()
        h = batch.edge_head[order].cpu().numpy(); r_ = rel[order].cpu().numpy(); t = batch.edge_tail[order].cpu().numpy()
        sc = score[order].detach().cpu().numpy()
        out = []
        starts = np.searchsorted(blk_s, np.arange(B + 1))
        for b in range(B):
            sl = slice(starts[b], starts[b + 1])
            kk = keep_s[sl]
            n = min(int(kk.sum()), k)
            out.append((np.stack([h[sl][:n], r_[sl][:n], t[sl][:n]], 1).astype(np.int64), sc[sl][:n]))
        return out

    def explain(self, triples, init_value, *, epochs=5, learning_rate=1e-3, threshold=0.2, structure_target=None,
                shared_optimizer=False, batch_size=512, k=10):
        """GNNExplainer (structure_target=None, epochs=5, threshold=0.2) or IDDGCN_explain (structure_target=
        [0.4,0.4,0.1,0.1], epochs=10, threshold=0.15).  init_value: (N,N) array — the mask initialisation shared by
        all relations (GnnExplainer.py:123-134).  Returns [(triples [k',3] (h,r,t), scores [k'], loss)] per input."""
        triples = np.asarray(triples, np.int64).reshape(-1, 3)
        init = torch.as_tensor(np.asarray(init_value, np.float32), device=self.device)
        results = []
        if shared_optimizer:
            state = _DenseAdam(self, init)
            batch_size = 1
        import time as _time

        def _tick(name, t0):
            if self.timings is not None:
                torch.cuda.synchronize(self.device)
                self.timings[name] = self.timings.get(name, 0.0) + _time.perf_counter() - t0
            return _time.perf_counter()

        for s in range(0, len(triples), batch_size):
            t0 = _time.perf_counter()
            chunk = triples[s:s + batch_size]
            n_real = len(chunk)
            if not shared_optimizer and n_real < batch_size and batch_size in self._replicas:
                # pad the last batch to the size whose replica model already exists (padding results are dropped)
                chunk = np.concatenate([chunk, np.repeat(chunk[:1], batch_size - n_real, axis=0)], 0)
            batch = _Batch(self, chunk)
            logits = init[batch.edge_head, batch.edge_tail].contiguous()
            t0 = _tick("batch_assembly", t0)
            if not shared_optimizer:
                state = _EdgeAdam(self, batch.graph.nnz)
            _, loss = self._optimise(batch, logits, epochs, learning_rate, structure_target, state)
            t0 = _tick("mask_optimisation", t0)
            if shared_optimizer:
                logits = state.current_logits(batch)
            top = self._topk(batch, logits, threshold, k)
            lv = loss.cpu().numpy()
            results += [(a, b, float(lv[i])) for i, (a, b) in enumerate(top)][:n_real]
            batch.close()
            t0 = _tick("top_k", t0)
            if shared_optimizer:
                state.reset_masks()
        return results

    # -- explaiNE ---------------------------------------------------------------------------------------
    def explaine(self, triples, k=10):
        """explaiNE.py:12-32,84-96: top-k edges of the FULL graph by d(pred)/d(adjacency value), per triple."""
        triples = np.asarray(triples, np.int64).reshape(-1, 3)
        g, m = self.graph, self.src
        d = g.device_arrays()
        heads = d["csr_row"].cpu().numpy().astype(np.int64)
        tails = d["csr_col"].cpu().numpy().astype(np.int64)
        rels = np.repeat(np.arange(self.R), g.nnz_per_rel)
        ones = torch.ones(g.nnz, dtype=torch.float32, device=self.device)
        one = torch.ones(1, dtype=torch.float32, device=self.device)
        out = []
        for h, r, t in triples:
            plan = m._plan(np.array([h]), np.array([r]), np.array([t]))
            pred = m.forward_values(g, plan, ones, 1)
            dv = m.backward_values(g, plan, one, ones, 1, want_dvals=True).cpu().numpy()
            order = np.argsort(-dv, kind="stable")[:k]                   # sorted(..., reverse=True)[:top_k], stable
            out.append((np.stack([heads[order], rels[order], tails[order]], 1), dv[order], float(pred.item())))
        return out


class _EdgeAdam:
    """Fresh Keras-Adam state over the per-edge logits of one batch."""

    def __init__(self, ex: Explainer, n: int):
        self.ex = ex
        self.m = torch.zeros(max(n, 1), dtype=torch.float32, device=ex.device)
        self.v = torch.zeros(max(n, 1), dtype=torch.float32, device=ex.device)
        self.t = 0

    def apply(self, logits, grad, batch, lr):
        self.t += 1
        check(self.ex._lib.iddgcn_adam_step(ptr(logits), ptr(grad), ptr(self.m), ptr(self.v), logits.numel(), lr, 0.9,
                                            0.999, 1e-7, self.t, _stream(self.ex.device)))


class _DenseAdam:
    """The reference's shared optimizer: dense (R,N,N) masks and Adam slots that persist across triples."""

    def __init__(self, ex: Explainer, init):
        self.ex = ex
        R, N = ex.R, ex.N
        self.init = init
        self.masks = init.unsqueeze(0).repeat(R, 1, 1).contiguous()
        self.m = torch.zeros_like(self.masks)
        self.v = torch.zeros_like(self.masks)
        self.grad = torch.zeros_like(self.masks)
        self.t = 0

    def apply(self, logits, grad, batch, lr):
        idx = (batch.edge_rel, batch.edge_head, batch.edge_tail)
        self.grad.zero_()
        self.grad[idx] = grad
        self.t += 1
        check(self.ex._lib.iddgcn_adam_step(ptr(self.masks), ptr(self.grad), ptr(self.m), ptr(self.v),
                                            self.masks.numel(), lr, 0.9, 0.999, 1e-7, self.t, _stream(self.ex.device)))
        logits.copy_(self.masks[idx])

    def current_logits(self, batch):
        return self.masks[(batch.edge_rel, batch.edge_head, batch.edge_tail)].contiguous()

    def reset_masks(self):
        self.masks.copy_(self.init.unsqueeze(0).expand_as(self.masks))     # mask.assign(init_value), :72-73
