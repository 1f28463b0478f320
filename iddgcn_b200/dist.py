"""Destination-row partitioned relational SpMM over the GPUs of one NVSwitch box (BASELINE configs[3]).

Rank p owns destination rows [p·B, (p+1)·B) of every relation (its Graph is built with row_block=...), the matching
rows of S / dS, and the same block of the features X.  One exchange per direction:

    forward :  all-gather the feature blocks        -> X_full          -> S_local  = A_local · X_full
    backward:  dX_partial = Σ_r A_local,rᵀ · dS_local -> reduce-scatter -> dX_local

The feature dimension is cut into `panels` column panels (layout [panel][row][D/panels], each panel a compact
matrix) so that NCCL moves panel c+1 over NVLink while the gather kernel works on panel c: both collectives are
issued asynchronously on NCCL's stream and only their completion events are awaited on the compute stream.
`torch.distributed` is plumbing; the SpMM kernels are the hand-written ones behind the C ABI.

The local compute is injected (`fwd_fn` / `bwd_fn`) so that the partition / collective logic can be exercised by
world_size-2 gloo tests on the CPU; the default functions call the CUDA kernels and there is no CPU fallback here.
"""
from __future__ import annotations

import time

import torch
import torch.distributed as dist


def row_block(rank: int, world: int, n_total: int) -> tuple[int, int, int]:
    """(lo, hi, block) — equal contiguous destination-row blocks, the last one possibly short."""
    block = (n_total + world - 1) // world
    lo = min(rank * block, n_total)
    return lo, min(lo + block, n_total), block


def panelize(x: torch.Tensor, panels: int) -> torch.Tensor:
    """[..., n, D] row-major -> [panels, ..., n, D/panels] (each panel compact)."""
    *lead, n, D = x.shape
    if D % panels:
        raise ValueError(f"D={D} is not divisible by panels={panels}")
    return x.reshape(*lead, n, panels, D // panels).movedim(-2, 0).contiguous()


def unpanelize(xp: torch.Tensor) -> torch.Tensor:
    """Inverse of panelize."""
    panels = xp.shape[0]
    y = xp.movedim(0, -2)
    *lead, n, p, dc = y.shape
    return y.reshape(*lead, n, p * dc).contiguous()


def _all_gather(out: torch.Tensor, inp: torch.Tensor, async_op: bool, group=None):
    return dist.all_gather_into_tensor(out, inp, group=group, async_op=async_op)


def _reduce_scatter(out: torch.Tensor, inp: torch.Tensor, rank: int, async_op: bool, group=None):
    if dist.get_backend() == "nccl":
        return dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    # host-logic path for the gloo tests (no reduce-scatter there): all-reduce, keep this rank's block
    dist.all_reduce(inp, op=dist.ReduceOp.SUM, group=group)
    n = out.shape[0]
    out.copy_(inp[rank * n:(rank + 1) * n])
    return None


class _NvlsBuffers:
    """Symmetric (multicast-mapped) buffers + the hand-written multimem kernels of csrc/nvls.cu.

    torch.distributed._symmetric_memory is the plumbing (cuMulticast* object, handle exchange, signal-pad barriers);
    the data movement is ours: `iddgcn_mc_store` (all-gather leg, multimem.st) and `iddgcn_mc_reduce_load`
    (reduce-scatter leg, multimem.ld_reduce.add)."""

    def __init__(self, shape, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.full = symm_mem.empty(*shape, dtype=torch.float32, device=device)
        self.partial = symm_mem.empty(*shape, dtype=torch.float32, device=device)
        self.h_full = symm_mem.rendezvous(self.full, group)
        self.h_part = symm_mem.rendezvous(self.partial, group)
        self.mc_full = int(self.h_full.multicast_ptr)
        self.mc_part = int(self.h_part.multicast_ptr)
        if not self.mc_full or not self.mc_part:
            raise RuntimeError("NVLS multicast is not available on this system")
        self.s_ag = torch.cuda.Stream(device=device, priority=-1)
        self.s_rs = torch.cuda.Stream(device=device, priority=-1)


class PartitionedSpMM:
    def __init__(self, graph, D: int, world: int, rank: int, block: int, *, panels: int = 4, fwd_fn=None,
                 bwd_fn=None, device=None, ag_group=None, rs_group=None, exchange: str = "nccl", mc_ctas: int = 32):
        if D % panels or (D // panels) % 4:
            raise ValueError(f"D={D} must split into {panels} panels whose width is a multiple of 4")
        self.g, self.D, self.world, self.rank, self.block, self.panels = graph, D, world, rank, block, panels
        self.Dc = D // panels
        # two communicators (two NCCL streams): the all-gather is receive-heavy and the reduce-scatter send-heavy, so
        # running them concurrently keeps both NVLink directions busy
        self.ag_group, self.rs_group = ag_group, rs_group
        self.R = graph.R
        self.n_rows = graph.n_rows
        self.N_pad = block * world            # all-gather needs equal blocks; the last block may carry padding rows
        if fwd_fn is None or bwd_fn is None:
            from . import ops                 # CUDA kernels; raises on a machine without a GPU
            fwd_fn = fwd_fn or (lambda g, X, out: ops.spmm_fwd(g, X, out=out))
            bwd_fn = bwd_fn or (lambda g, dS, out: ops.spmm_bwd_x(g, dS, out=out))
        self.fwd_fn, self.bwd_fn = fwd_fn, bwd_fn
        dev = device if device is not None else graph.device
        f = dict(dtype=torch.float32, device=dev)
        self.Xp_local = torch.zeros((panels, block, self.Dc), **f)
        self.exchange, self.mc_ctas, self.nvls = exchange, mc_ctas, None
        if exchange == "nvls":
            # NVSwitch multicast exchange: both big buffers must be symmetric (same offset on every rank)
            self.nvls = _NvlsBuffers((panels, self.N_pad, self.Dc), dev, dist.group.WORLD)
            self.Xp_full, self.dXp_partial = self.nvls.full, self.nvls.partial
        else:
            self.Xp_full = torch.empty((panels, self.N_pad, self.Dc), **f)
            self.dXp_partial = torch.empty((panels, self.N_pad, self.Dc), **f)
        self.dXp_local = torch.empty((panels, block, self.Dc), **f)
        self.Sp = torch.empty((panels, self.R, self.n_rows, self.Dc), **f)
        self.dSp = torch.empty((panels, self.R, self.n_rows, self.Dc), **f)
        if self.N_pad != graph.N and graph.N > self.N_pad:
            raise ValueError("block * world must cover all source nodes")

    # -- layout helpers -----------------------------------------------------------------------------
    def load_inputs(self, X_local: torch.Tensor, dS_local: torch.Tensor):
        """Row-major rank-local inputs -> panel layout (done once per step's inputs, outside the exchange)."""
        self.Xp_local[:, :self.n_rows].copy_(panelize(X_local, self.panels))
        self.dSp.copy_(panelize(dS_local, self.panels))

    def outputs(self):
        """(S_local [R,n_rows,D], dX_local [n_rows,D]) back in row-major layout."""
        return unpanelize(self.Sp), unpanelize(self.dXp_local[:, :self.n_rows])

    # -- one exchange + compute step -------------------------------------------------------------------
    def run_panels(self):
        """One step: backward panels feed the reduce-scatter while the all-gather feeds the forward panels.

        Issue order on the compute stream is B0 F0 B1 F1 ...: backward panel c needs no exchange (dS is local), so its
        kernel and the all-gather of panel c start together; its result goes straight into reduce-scatter c on the
        second communicator while forward panel c consumes all-gather c."""
        if self.nvls is not None:
            return self._run_panels_nvls()
        P = self.panels
        gathers = [_all_gather(self.Xp_full[c], self.Xp_local[c], True, self.ag_group) for c in range(P)]
        scatters = []
        for c in range(P):
            part = self.dXp_partial[c]
            self.bwd_fn(self.g, self.dSp[c], part[: self.g.N])
            if self.N_pad > self.g.N:
                part[self.g.N:].zero_()
            scatters.append(_reduce_scatter(self.dXp_local[c], part, self.rank, True, self.rs_group))
            if gathers[c] is not None:
                gathers[c].wait()
            self.fwd_fn(self.g, self.Xp_full[c][: self.g.N], self.Sp[c])
        for w in scatters:
            if w is not None:
                w.wait()

    def _run_panels_nvls(self):
        """Same step with the exchange done by our multimem kernels on two side streams.

        all-gather leg  : every rank stores its block of panel c once through the multicast address (the switch
                          replicates it into every rank's Xp_full) -> barrier -> forward panel c may read it;
        reduce-scatter  : after backward panel c on every rank (barrier), each rank loads ITS block of dXp_partial[c]
                          through the multicast address with the in-switch sum.
        The two legs load opposite link directions and run concurrently with the kernels."""
        from ._lib import check, load
        lib, nv, P = load(), self.nvls, self.panels
        main = torch.cuda.current_stream(self.Xp_local.device)
        blk = self.block * self.Dc                       # floats per (panel, rank) block
        T_MS = 20000
        for s_ in (nv.s_ag, nv.s_rs):
            main.wait_stream(s_)
        nv.h_full.barrier(channel=2, timeout_ms=T_MS)    # every rank is done with the previous step's buffers
        nv.s_ag.wait_stream(main)
        nv.s_rs.wait_stream(main)
        ev_ag = []
        with torch.cuda.stream(nv.s_ag):
            for c in range(P):
                off = (c * self.N_pad * self.Dc + self.rank * blk) * 4
                check(lib.iddgcn_mc_store(nv.mc_full + off, self.Xp_local[c].data_ptr(), blk, self.mc_ctas,
                                          nv.s_ag.cuda_stream))
                nv.h_full.barrier(channel=0, timeout_ms=T_MS)
                ev_ag.append(nv.s_ag.record_event())
        ev_rs = []
        for c in range(P):
            part = self.dXp_partial[c]
            self.bwd_fn(self.g, self.dSp[c], part[: self.g.N])
            if self.N_pad > self.g.N:
                part[self.g.N:].zero_()
            ev_b = main.record_event()
            with torch.cuda.stream(nv.s_rs):
                nv.s_rs.wait_event(ev_b)
                nv.h_part.barrier(channel=0, timeout_ms=T_MS)    # every rank's partial panel c is complete
                off = (c * self.N_pad * self.Dc + self.rank * blk) * 4
                check(lib.iddgcn_mc_reduce_load(nv.mc_part + off, self.dXp_local[c].data_ptr(), blk, self.mc_ctas,
                                                nv.s_rs.cuda_stream))
                ev_rs.append(nv.s_rs.record_event())
            main.wait_event(ev_ag[c])
            self.fwd_fn(self.g, self.Xp_full[c][: self.g.N], self.Sp[c])
        for e in ev_rs:
            main.wait_event(e)

    def step(self, X_local, dS_local, S_out=None):
        """bench.py's step: inputs are already resident in panel layout after the first call."""
        if not getattr(self, "_loaded", False):
            self.load_inputs(X_local, dS_local)
            self._loaded = True
        self.run_panels()

    # -- measurement helpers used by bench.py ---------------------------------------------------------
    def kernel_times(self, X_local, dS_local, S_out, iters=5):
        """Per-rank compute-only times (no collectives): all panels fwd, all panels bwd."""
        out = {"fwd": [], "bwd": []}
        for _ in range(iters):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            for p in range(self.panels):
                self.fwd_fn(self.g, self.Xp_full[p][: self.g.N], self.Sp[p])
            b.record()
            for p in range(self.panels):
                self.bwd_fn(self.g, self.dSp[p], self.dXp_partial[p][: self.g.N])
            c.record()
            torch.cuda.synchronize()
            out["fwd"].append(a.elapsed_time(b))
            out["bwd"].append(b.elapsed_time(c))
        return out

    def comm_times(self, iters=5):
        """Per-rank collective-only times: all panels all-gather, all panels reduce-scatter."""
        out = {"all_gather": [], "reduce_scatter": []}
        for _ in range(iters):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            dist.barrier()
            a.record()
            for p in range(self.panels):
                _all_gather(self.Xp_full[p], self.Xp_local[p], False, self.ag_group)
            b.record()
            for p in range(self.panels):
                _reduce_scatter(self.dXp_local[p], self.dXp_partial[p], self.rank, False, self.rs_group)
            c.record()
            torch.cuda.synchronize()
            out["all_gather"].append(a.elapsed_time(b))
            out["reduce_scatter"].append(b.elapsed_time(c))
        return out

    def e2e(self, args, nnz_global):
        """Same step with HOST buffers: pinned X/dS blocks up, S/dX blocks down, every step; max over ranks."""
        Xh = self.Xp_local.cpu().pin_memory()
        dSh = self.dSp.cpu().pin_memory()
        Sh = torch.empty(self.Sp.shape).pin_memory()
        dXh = torch.empty(self.dXp_local.shape).pin_memory()
        k = max(3, min(args.steps, 5))

        def one():
            self.Xp_local.copy_(Xh, non_blocking=True)
            self.dSp.copy_(dSh, non_blocking=True)
            self.run_panels()
            Sh.copy_(self.Sp, non_blocking=True)
            dXh.copy_(self.dXp_local, non_blocking=True)

        one()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            one()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / k], dtype=torch.float64, device=self.Xp_local.device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        return {"value": nnz_global / dt, "unit": "edges/s",
                "h2d_bytes_per_step": (Xh.numel() + dSh.numel()) * 4 * self.world,
                "d2h_bytes_per_step": (Sh.numel() + dXh.numel()) * 4 * self.world, "ms_per_step": dt * 1e3, "steps": k,
                "api": "PartitionedSpMM.run_panels with pinned host blocks in/out on every rank"}
