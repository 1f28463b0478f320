"""`IDDGCN_Layer` and `DistMult` with the reference's constructor and call signatures
(prediction/IDDGCN.py:16-79, 82-109), as torch.nn.Modules whose forward and backward run in libiddgcn_b200.

    layer = IDDGCN_Layer(num_entities, num_relations, output_dim, seed)
    head_out, tail_out = layer([embeddings, head_idx, head_e, tail_idx, tail_e, adj_list])
    score = DistMult(num_relations, seed)([head_e, rel_idx, tail_e])          # (1, T)

Gradients flow (through torch's tape, computed by the CUDA kernels) to: embeddings, head_e, tail_e, every layer
weight, and the adjacency values when `adj_list` carries them (explainers).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import VALS_ONES, check, load, ptr
from .graph import collect_edge_values
from .plan import Plan


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _PlanCache:
    def __init__(self, N, R):
        self.N, self.R = N, R
        self._c = {}

    def get(self, head_idx, rel_idx, tail_idx) -> Plan:
        def key_of(x):
            if x is None:
                return None
            if isinstance(x, torch.Tensor) and x.is_cuda:
                return ("d", x.data_ptr(), x.numel(), x._version)
            a = x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
            return ("h", a.size, hash(np.ascontiguousarray(a).tobytes()))
        key = (key_of(head_idx), key_of(rel_idx), key_of(tail_idx))
        ent = self._c.get(key)
        p = ent[0] if ent is not None else None
        if p is None:
            if len(self._c) >= 32:
                self._c.pop(next(iter(self._c)))[0].close()
            ref = head_idx if head_idx is not None else rel_idx
            T = int(ref.numel() if isinstance(ref, torch.Tensor) else np.asarray(ref).size)
            dev = ref.device if isinstance(ref, torch.Tensor) and ref.is_cuda else None
            zeros = torch.zeros(T, dtype=torch.int64, device=dev) if dev is not None else np.zeros(T, np.int64)
            p = Plan(head_idx if head_idx is not None else zeros, rel_idx if rel_idx is not None else zeros,
                     tail_idx if tail_idx is not None else zeros, self.N, self.R,
                     device=dev)
            # device tensors are keyed by address: keep them alive with the entry so the address cannot be recycled
            self._c[key] = (p, [x for x in (head_idx, rel_idx, tail_idx) if isinstance(x, torch.Tensor) and x.is_cuda])
        return p


class _LayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, head_e, tail_e, rk, sk, wa, ba, vals, graph, plan, mode):
        lib = load()
        f = lambda t: t.detach().to(torch.float32).contiguous()
        emb, rk, sk, wa, ba = f(emb), f(rk), f(sk), f(wa), f(ba)
        vals_c = f(vals) if vals is not None else None
        R, N, D, T = graph.R, graph.N, emb.shape[1], plan.T
        st = _stream(emb)
        S = ops.spmm_fwd(graph, emb, vals_c, mode)                                        # IDDGCN.py:70
        Z = torch.empty_like(S)
        check(lib.iddgcn_relation_transform(ptr(S), ptr(rk), R, N, D, ptr(Z), st))         # IDDGCN.py:76-77 (hoisted)
        HT_in = torch.stack([f(head_e), f(tail_e)], 0).contiguous()
        HT_out = torch.empty_like(HT_in)
        alpha = torch.empty((T, R), dtype=torch.float32, device=emb.device)
        check(lib.iddgcn_layer_fwd(plan.handle, None, ptr(HT_in[0]), ptr(HT_in[1]), ptr(Z), ptr(sk), ptr(wa), ptr(ba),
                                   D, ptr(HT_out[0]), ptr(HT_out[1]), ptr(alpha), st))
        ctx.graph, ctx.plan, ctx.mode = graph, plan, mode
        ctx.has_vals = vals is not None
        ctx.save_for_backward(emb, rk, sk, wa, S, Z, HT_in, HT_out, alpha, vals_c if vals_c is not None else emb.new_empty(0))
        return HT_out[0], HT_out[1]

    @staticmethod
    def backward(ctx, dH, dT):
        lib = load()
        emb, rk, sk, wa, S, Z, HT_in, HT_out, alpha, vals = ctx.saved_tensors
        vals = vals if ctx.has_vals else None
        graph, plan, mode = ctx.graph, ctx.plan, ctx.mode
        R, N, D = graph.R, graph.N, emb.shape[1]
        st = _stream(emb)
        dHT_out = torch.stack([dH.to(torch.float32), dT.to(torch.float32)], 0).contiguous()
        dHT_in = torch.empty_like(dHT_out)
        dZ = torch.empty_like(Z)
        dsk, dwa = torch.zeros_like(sk), torch.zeros_like(wa)
        dba = torch.zeros(R, dtype=torch.float32, device=emb.device)
        check(lib.iddgcn_layer_bwd(plan.handle, ptr(HT_in), ptr(HT_out), ptr(alpha), ptr(Z), ptr(sk), ptr(wa), D,
                                   ptr(dHT_out), ptr(dHT_in), ptr(dZ), 0, ptr(dsk), ptr(dwa), ptr(dba), st))
        drk = torch.zeros_like(rk)
        dS = torch.empty_like(S)
        check(lib.iddgcn_relation_transform_bwd(ptr(S), ptr(rk), ptr(dZ), R, N, D, ptr(drk), ptr(dS), 0, st))
        demb = ops.spmm_bwd_x(graph, dS, vals, mode) if ctx.needs_input_grad[0] else None
        dvals = ops.spmm_bwd_vals(graph, dS, emb, vals, mode) if (vals is not None and ctx.needs_input_grad[7]) else None
        return demb, dHT_in[0], dHT_in[1], drk, dsk, dwa, dba, dvals, None, None, None


class IDDGCN_Layer(torch.nn.Module):
    def __init__(self, num_entities, num_relations, output_dim, seed, **kwargs):
        super().__init__()
        self.num_entities, self.num_relations, self.output_dim, self.seed = num_entities, num_relations, output_dim, seed
        g = np.random.default_rng(seed)
        R, D = num_relations, output_dim
        lim = np.sqrt(6.0 / (D + R))
        mk = lambda a: torch.nn.Parameter(torch.tensor(a, dtype=torch.float32, device="cuda"))
        # creation order = weight-file slot order (IDDGCN.py:25-58)
        self.relation_kernel = mk(g.standard_normal((R, D, D)))
        self.self_kernel = mk(g.standard_normal((D, D)))
        self.relation_weights = mk(g.uniform(-0.05, 0.05, (R,)))      # created, never used in call (IDDGCN.py:39-44)
        self.W_alpha = mk(g.uniform(-lim, lim, (D, R)))
        self.b_alpha = mk(np.zeros(R))
        self._plans = _PlanCache(num_entities, num_relations)

    def forward(self, inputs):
        embeddings, head_idx, head_e, tail_idx, tail_e, *adj = inputs             # IDDGCN.py:61
        adj_list = adj[0] if len(adj) == 1 and isinstance(adj[0], (list, tuple)) else adj
        graph, vals, mode = collect_edge_values(list(adj_list))
        plan = self._plans.get(head_idx, None, tail_idx)
        return _LayerFn.apply(embeddings, head_e, tail_e, self.relation_kernel, self.self_kernel, self.W_alpha,
                              self.b_alpha, vals, graph, plan, mode)

    call = forward


class _DistMultFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head_e, tail_e, kernel, plan):
        lib = load()
        HT = torch.stack([head_e.detach().to(torch.float32), tail_e.detach().to(torch.float32)], 0).contiguous()
        k = kernel.detach().to(torch.float32).contiguous()
        y = torch.empty(plan.T, dtype=torch.float32, device=HT.device)
        check(lib.iddgcn_distmult_fwd(plan.handle, ptr(HT), ptr(k), HT.shape[2], ptr(y), _stream(HT)))
        ctx.plan = plan
        ctx.save_for_backward(HT, k, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = load()
        HT, k, y = ctx.saved_tensors
        dHT = torch.empty_like(HT)
        dk = torch.zeros_like(k)
        dyc = dy.to(torch.float32).contiguous()
        check(lib.iddgcn_distmult_bwd(ctx.plan.handle, ptr(HT), ptr(k), ptr(y), ptr(dyc), HT.shape[2], ptr(dHT),
                                      ptr(dk), _stream(HT)))
        return dHT[0], dHT[1], dk, None


class DistMult(torch.nn.Module):
    def __init__(self, num_relations, seed, embedding_dim=None, **kwargs):
        super().__init__()
        self.num_relations, self.seed = num_relations, seed
        self.kernel = None
        self._plans = None
        if embedding_dim is not None:
            self.build(embedding_dim)

    def build(self, embedding_dim):
        """Keras builds lazily from the input shape (IDDGCN.py:89-101); pass embedding_dim to build eagerly."""
        g = np.random.default_rng(self.seed)
        self.kernel = torch.nn.Parameter(
            torch.tensor(g.standard_normal((self.num_relations, embedding_dim)), dtype=torch.float32, device="cuda"))

    def forward(self, inputs):
        head_e, rel_idx, tail_e = inputs                                           # IDDGCN.py:104
        if self.kernel is None:
            self.build(head_e.shape[-1])
        if self._plans is None:
            self._plans = _PlanCache(1, self.num_relations)
        plan = self._plans.get(None, rel_idx, None)
        return _DistMultFn.apply(head_e, tail_e, self.kernel, plan).unsqueeze(0)   # IDDGCN.py:109

    call = forward
