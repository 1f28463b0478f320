"""Grid-partitioned relational SpMM over the GPUs of one NVSwitch box (BASELINE configs[3]).

The ranks form a P_r x P_c grid.  Rank (pr, pc) owns

    * destination rows [pr·B, (pr+1)·B) of every relation (its Graph is built with row_block=...),
    * feature columns  [pc·Dc, (pc+1)·Dc),  Dc = D / P_c,

i.e. the blocks X[rows_pr, cols_pc], dS[:, rows_pr, cols_pc] and the outputs S[:, rows_pr, cols_pc],
dX[rows_pr, cols_pc].  An SpMM is independent per feature column, so ranks that differ in pc never talk; the P_r
ranks of one column group exchange rows:

    forward :  all-gather the X blocks of the group        -> X_full[N, Dc]   -> S_local  = A_local · X_full
    backward:  dX_partial = Σ_r A_local,rᵀ · dS_local [N, Dc]  -> reduce-scatter over the group -> dX_local

Per rank the exchange moves (P_r-1)/P_r · N · 4·D/P_c bytes each way: 8.96 GB at 8x1 on the 20M-node graph, 1.28 GB at
2x4, nothing at 1xP.  P_c = 1 is the plain destination-row partition; P_r = 1 replicates the graph and splits only
the feature columns.

Exchange "p2p" (default, CUDA): no collective call on the data path.
    * reduce-scatter: FUSED into the backward kernel.  The kernel is launched per owner block; for a peer's block its
      output pointer is that peer's stage buffer (NVLink peer stores, the partial rows travel while the kernel runs);
      for the rank's own block the kernel's epilogue adds the partials the peers pushed (fixed order -> reproducible).
    * all-gather: each rank's copy engines push its X block into every peer's X_full on a side stream.
    * ordering: 32-bit epoch flags in symmetric memory (iddgcn_peer_signal / iddgcn_peer_wait), no host barrier.
  torch.distributed._symmetric_memory only allocates and maps the buffers (plumbing).
Exchange "nccl": all_gather_into_tensor / reduce_scatter_tensor inside the column group (the baseline; also the path
the world_size-2 gloo tests exercise on the CPU with injected compute).

The local compute is injected (`fwd_fn` / `bwd_fn`) only by those CPU tests; the default functions call the CUDA kernels
and there is no CPU fallback here.
"""
from __future__ import annotations

import ctypes as C
import time

import torch
import torch.distributed as dist

# Default grids, from per-rank kernel times measured for every layout on the 20M-node graph (profiles/
# r02_grid_shapes_*.json) and the 2-GPU runs (profiles/r02_bench_c4_2gpu_*.json): narrow rows gather less efficiently
# (D/P_c = 128 / 64 / 32 / 16 floats run at 1.03 / 1.00 / 0.86 / 0.5 of the HBM roofline) while every extra row block
# adds NVLink traffic, so columns are split first and rows only once the rows get too narrow (16 floats).
DEFAULT_GRIDS = {1: (1, 1), 2: (1, 2), 4: (1, 4), 8: (2, 4)}
K_DONE, K_AG, K_PART = 0, 1, 2          # flag kinds
SLOTS = 8                               # flag slots per kind (max ranks in a column group)


def row_block(rank: int, world: int, n_total: int) -> tuple[int, int, int]:
    """(lo, hi, block) — equal contiguous destination-row blocks, the last ones possibly short or empty."""
    block = (n_total + world - 1) // world
    lo = min(rank * block, n_total)
    return lo, min(lo + block, n_total), block


def parse_grid(spec, world: int, D: int) -> tuple[int, int]:
    """'RxC' / (R, C) / None -> (P_r, P_c) with P_r·P_c == world and D divisible into P_c column groups of 4k floats."""
    if spec is None or spec == "auto":
        pr, pc = DEFAULT_GRIDS.get(world, (world, 1))
        while pc > 1 and D % (4 * pc):
            pc //= 2
            pr = world // pc
    elif isinstance(spec, str):
        pr, pc = (int(x) for x in spec.lower().split("x"))
    else:
        pr, pc = int(spec[0]), int(spec[1])
    if pr * pc != world:
        raise ValueError(f"grid {pr}x{pc} does not match world size {world}")
    if D % pc or (D // pc) % 4:
        raise ValueError(f"D={D} does not split into {pc} column groups whose width is a multiple of 4")
    return pr, pc


def grid_coords(rank: int, grid: tuple[int, int]) -> tuple[int, int]:
    """rank -> (pr, pc); ranks of one column group are pc, pc+P_c, pc+2·P_c, ..."""
    return rank // grid[1], rank % grid[1]


def _reduce_scatter(out, inp, index, group):
    if dist.get_backend() == "nccl":
        return dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group, async_op=True)
    # host-logic path for the gloo tests (no reduce-scatter there): all-reduce, keep this rank's block
    dist.all_reduce(inp, op=dist.ReduceOp.SUM, group=group)
    n = out.shape[0]
    out.copy_(inp[index * n:(index + 1) * n])
    return None


class GridSpMM:
    def __init__(self, graph, D: int, world: int, rank: int, grid=None, *, exchange: str = "p2p", fwd_fn=None,
                 bwd_fn=None, device=None, timeout_ms: int = 20000, partials: str = "stores", order: str = "ABF"):
        self.g, self.D, self.world, self.rank = graph, D, world, rank
        self.P_r, self.P_c = parse_grid(grid, world, D)
        self.pr, self.pc = grid_coords(rank, (self.P_r, self.P_c))
        self.Dc = D // self.P_c
        self.R, self.N = graph.R, graph.N
        self.lo, self.hi, self.block = row_block(self.pr, self.P_r, self.N)
        self.n_rows = self.hi - self.lo
        if graph.n_rows != self.n_rows:
            raise ValueError(f"graph holds {graph.n_rows} rows, grid position ({self.pr},{self.pc}) owns {self.n_rows}")
        self.N_pad = self.block * self.P_r          # equal blocks for the collectives; the tail may be padding
        self.group_ranks = [q * self.P_c + self.pc for q in range(self.P_r)]
        self.peers = [q for q in range(self.P_r) if q != self.pr]          # group indices of the other ranks
        self.injected = fwd_fn is not None or bwd_fn is not None
        self.exchange = exchange if self.P_r > 1 else "none"
        self.timeout_ms = timeout_ms
        if partials not in ("stores", "dma") or order not in ("AFB", "ABF"):
            raise ValueError("partials must be 'stores' or 'dma', order 'AFB' or 'ABF'")
        # how the partial dX rows of peer-owned nodes travel: "stores" = the backward kernel writes them straight into
        # the owner's stage buffer (fused, NVLink stores from the SMs); "dma" = the kernel writes them locally and the
        # copy engines push them while the forward kernel runs.  order: A = peers' node blocks (backward), F = forward,
        # B = own node block (backward + staged partials).  Measured on 2 GPUs, 2x1 grid, 20M-node graph (kernels alone
        # 18.2 ms): stores/ABF 23.7 ms (= the sum of its NVLink-bound phases), dma/AFB 24.1, dma/ABF 27.5, stores/AFB
        # 28.9 (partial stores and the all-gather push share the links while only A can run), NCCL 34.3.
        self.partials, self.order = partials, order
        dev = device if device is not None else graph.device
        self.dev = dev
        f = dict(dtype=torch.float32, device=dev)
        self.col_group = None
        if self.P_r > 1 and (self.exchange == "nccl" or self.injected):
            # every rank creates every column group (new_group is collective over the world)
            for pc in range(self.P_c):
                grp = dist.new_group([q * self.P_c + pc for q in range(self.P_r)])
                if pc == self.pc:
                    self.col_group = grp
        if self.injected:
            self.fwd_fn, self.bwd_fn = fwd_fn, bwd_fn
            self.exchange = "nccl" if self.P_r > 1 else "none"
        else:
            from . import _lib                    # CUDA library; raises on a machine without a GPU / nvcc
            self._lib = _lib
            self.lib = _lib.load()
        self.epoch = 0
        self.stage = None
        if self.exchange == "p2p":
            self._init_p2p()
        else:
            self.X_full = torch.zeros((self.N_pad, self.Dc), **f)
            if self.exchange == "nccl":
                self.dX_partial = torch.zeros((self.N_pad, self.Dc), **f)
        self.S = torch.empty((self.R, self.n_rows, self.Dc), **f)
        self.dS = torch.empty((self.R, self.n_rows, self.Dc), **f)
        self.dX = torch.empty((max(self.block, 1), self.Dc), **f)[: self.block]
        if not self.injected:
            self.s_side = torch.cuda.Stream(device=dev, priority=-1)
            self._lib.check(self.lib.iddgcn_spmm_prepare(graph.handle, self.Dc,
                                                         torch.cuda.current_stream(dev).cuda_stream))

    # -- symmetric buffers of the fused exchange --------------------------------------------------------
    def _init_p2p(self):
        import torch.distributed._symmetric_memory as symm_mem
        dev, f32 = self.dev, torch.float32
        nslot = max(self.P_r - 1, 1)
        self.X_full = symm_mem.empty(self.N_pad, self.Dc, dtype=f32, device=dev)
        self.stage = symm_mem.empty(nslot, self.block, self.Dc, dtype=f32, device=dev)
        self.flags = symm_mem.empty(3 * SLOTS, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.X_full.zero_()
        self.h_x = symm_mem.rendezvous(self.X_full, dist.group.WORLD)
        self.h_stage = symm_mem.rendezvous(self.stage, dist.group.WORLD)
        self.h_flags = symm_mem.rendezvous(self.flags, dist.group.WORLD)
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.part_local = (torch.empty((nslot, self.block, self.Dc), dtype=f32, device=dev)
                           if self.partials == "dma" else None)
        torch.cuda.synchronize(dev)
        dist.barrier()                              # every rank's flags are zero before anyone signals
        xp, sp, fp = self.h_x.buffer_ptrs, self.h_stage.buffer_ptrs, self.h_flags.buffer_ptrs
        blk_bytes = self.block * self.Dc * 4
        me = self.pr
        self.peer_x_dst = []                        # where my X block goes in each peer's X_full
        self.peer_stage_dst = []                    # my slot in each peer's stage buffer
        for q in self.peers:
            gr = self.group_ranks[q]
            self.peer_x_dst.append(int(xp[gr]) + me * blk_bytes)
            slot = me if me < q else me - 1
            self.peer_stage_dst.append(int(sp[gr]) + slot * blk_bytes)

        def slots(kind, remote):
            arr = (C.c_void_p * len(self.peers))()
            for i, q in enumerate(self.peers):
                if remote:                          # my slot on peer q
                    arr[i] = int(fp[self.group_ranks[q]]) + (kind * SLOTS + me) * 4
                else:                               # peer q's slot here
                    arr[i] = int(fp[self.rank]) + (kind * SLOTS + q) * 4
            return arr

        self.sig = {k: slots(k, True) for k in (K_DONE, K_AG, K_PART)}
        self.wts = {k: slots(k, False) for k in (K_DONE, K_AG, K_PART)}

    # -- views the caller fills / reads ---------------------------------------------------------------------
    def x_local_view(self) -> torch.Tensor:
        """This rank's block inside X_full: write the features here and the step needs no local copy."""
        return self.X_full[self.lo:self.hi]

    def load_inputs(self, X_local: torch.Tensor, dS_local: torch.Tensor):
        """Rank-local inputs: X_local [n_rows, Dc], dS_local [R, n_rows, Dc] (this rank's column slice)."""
        if X_local is not None:
            self.x_local_view().copy_(X_local)
        if dS_local is not None:
            self.dS.copy_(dS_local)

    def outputs(self):
        """(S_local [R, n_rows, Dc], dX_local [n_rows, Dc])."""
        return self.S, self.dX[: self.n_rows]

    # -- primitives ---------------------------------------------------------------------------------------
    def _st(self, stream=None):
        return (stream or torch.cuda.current_stream(self.dev)).cuda_stream

    def _signal(self, kind, stream=None):
        self._lib.check(self.lib.iddgcn_peer_signal(self.sig[kind], len(self.peers), self.epoch, self._st(stream)))

    def _wait(self, kind, value, stream=None):
        self._lib.check(self.lib.iddgcn_peer_wait(self.wts[kind], len(self.peers), value, self.timeout_ms,
                                                  self.timed_out.data_ptr(), self._st(stream)))

    def _fwd(self, lo=0, hi=None, stream=None):
        """S rows [lo, hi) of the flattened [R·n_rows] output."""
        total = self.R * self.n_rows
        hi = total if hi is None else hi
        if hi <= lo:
            return
        out = self.S.data_ptr() + lo * self.Dc * 4
        self._lib.check(self.lib.iddgcn_spmm_fwd_rows(self.g.handle, self.X_full.data_ptr(), self.Dc, None, 0, out, lo,
                                                      hi, self._st(stream)))

    def _bwd_rows(self, dst_ptr, node_lo, node_hi, addends=(), stream=None):
        if node_hi <= node_lo:
            return
        arr = (C.c_void_p * max(len(addends), 1))(*addends) if addends else None
        self._lib.check(self.lib.iddgcn_spmm_bwd_x_rows(self.g.handle, self.dS.data_ptr(), self.Dc, None, 0, dst_ptr,
                                                        node_lo, node_hi, arr, len(addends), self._st(stream)))

    def _push_partials(self, stream=None):
        """Backward over the node blocks owned by peers: written straight into their stage buffers (NVLink stores,
        partials == "stores") or into a local buffer the copy engines push afterwards (_dma_partials)."""
        for i, q in enumerate(self.peers):
            qlo, qhi, _ = row_block(q, self.P_r, self.N)
            dst = self.peer_stage_dst[i] if self.part_local is None else self.part_local[i].data_ptr()
            self._bwd_rows(dst, qlo, qhi, stream=stream)

    def _dma_partials(self, stream):
        for i, q in enumerate(self.peers):
            qlo, qhi, _ = row_block(q, self.P_r, self.N)
            self._lib.check(self.lib.iddgcn_copy_async(self.peer_stage_dst[i], self.part_local[i].data_ptr(),
                                                       (qhi - qlo) * self.Dc * 4, stream.cuda_stream))

    def _own_partial(self, a=0, b=None, stream=None):
        """Backward over (a chunk [a, b) of) this rank's own node block, folding in the staged peer partials."""
        b = self.n_rows if b is None else b
        off = a * self.Dc * 4
        adds = [self.stage.data_ptr() + (s * self.block * self.Dc) * 4 + off for s in range(len(self.peers))] \
            if self.exchange == "p2p" else []
        self._bwd_rows(self.dX.data_ptr() + off, self.lo + a, self.lo + b, adds, stream=stream)

    def _push_x(self, stream):
        """Copy engines: this rank's X block into every peer's X_full."""
        src = self.X_full.data_ptr() + self.lo * self.Dc * 4
        nbytes = self.n_rows * self.Dc * 4
        for dst in self.peer_x_dst:
            self._lib.check(self.lib.iddgcn_copy_async(dst, src, nbytes, stream.cuda_stream))

    # -- one exchange + compute step ------------------------------------------------------------------------
    def step(self, X_local=None, dS_local=None):
        """Inputs given here are loaded first (outside of what bench.py times they are already resident)."""
        if X_local is not None or dS_local is not None:
            self.load_inputs(X_local, dS_local)
        if self.injected:
            return self._step_injected()
        if self.exchange == "p2p":
            return self._step_p2p()
        if self.exchange == "nccl":
            return self._step_nccl()
        self._fwd()                                    # P_r == 1: nothing to exchange
        self._own_partial()

    def _step_p2p(self):
        main, side = torch.cuda.current_stream(self.dev), self.s_side
        tl = getattr(self, "timeline", None)           # measurement aid: (label, event) pairs of this step

        def mark(label, stream):
            if tl is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                tl.append((label, ev))

        self.epoch += 1
        e = self.epoch
        mark("start", main)
        self._wait(K_DONE, e - 1)                      # peers are done reading what I pushed in the previous step
        mark("peers done with previous step", main)
        ev0 = main.record_event()
        side.wait_event(ev0)
        if self.order == "AFB":
            # all-gather leg on the copy engines while the backward kernel produces (and, with partials == "stores",
            # streams over NVLink) the partials of the peers' node blocks
            self._push_x(side)
            self._signal(K_AG, side)
            mark("X block pushed (side)", side)
            self._push_partials()
            mark("A: peers' partials", main)
            if self.part_local is None:
                self._signal(K_PART)
            else:
                side.wait_event(main.record_event())
                self._dma_partials(side)               # rides the copy engines while the forward kernel runs
                self._signal(K_PART, side)
                mark("partials pushed (side)", side)
            self._wait(K_AG, e)                        # every peer's X block has landed here
            mark("X blocks arrived", main)
            self._fwd()
            mark("F: forward", main)
            self._wait(K_PART, e)                      # every peer's partials for my nodes have landed
            mark("partials arrived", main)
            self._own_partial()
            mark("B: own block + staged partials", main)
        else:
            # A B F: the partial stores have the links to themselves, then the all-gather push runs under B
            self._push_partials()
            mark("A: peers' partials", main)
            if self.part_local is None:
                self._signal(K_PART)
                side.wait_event(main.record_event())
            else:
                side.wait_event(main.record_event())
                self._dma_partials(side)
                self._signal(K_PART, side)
                mark("partials pushed (side)", side)
            self._push_x(side)
            self._signal(K_AG, side)
            mark("X block pushed (side)", side)
            self._wait(K_PART, e)
            mark("partials arrived", main)
            self._own_partial()
            mark("B: own block + staged partials", main)
            self._wait(K_AG, e)
            mark("X blocks arrived", main)
            self._fwd()
            mark("F: forward", main)
        self._signal(K_DONE)
        main.wait_stream(side)
        mark("end", main)

    def step_phases(self, iters=5):
        """Median completion time (ms since the step start, this rank) of every phase of the p2p step."""
        if self.exchange != "p2p":
            return None
        acc = {}
        for _ in range(iters):
            self.timeline = []
            self.step()
            torch.cuda.synchronize(self.dev)
            t0 = self.timeline[0][1]
            for label, ev in self.timeline[1:]:
                acc.setdefault(label, []).append(t0.elapsed_time(ev))
        self.timeline = None
        out = {}
        for k, v in acc.items():
            v.sort()
            out[k] = v[len(v) // 2]
        return out

    def _step_nccl(self):
        grp = self.col_group
        ag = dist.all_gather_into_tensor(self.X_full, self.X_full[self.pr * self.block:(self.pr + 1) * self.block],
                                         group=grp, async_op=True)
        self._bwd_rows(self.dX_partial.data_ptr(), 0, self.N)
        rs = _reduce_scatter(self.dX, self.dX_partial, self.pr, grp)
        ag.wait()
        self._fwd()
        if rs is not None:
            rs.wait()

    def _step_injected(self):
        """CPU host-logic path of the gloo tests: same partition arithmetic, compute from the injected functions."""
        if self.P_r > 1:
            mine = self.X_full[self.pr * self.block:(self.pr + 1) * self.block].clone()
            dist.all_gather_into_tensor(self.X_full, mine, group=self.col_group)
        self.fwd_fn(self.g, self.X_full[: self.N], self.S)
        part = torch.zeros((self.N_pad, self.Dc), dtype=torch.float32)
        self.bwd_fn(self.g, self.dS, part[: self.N])
        if self.P_r > 1:
            _reduce_scatter(self.dX, part, self.pr, self.col_group)
        else:
            self.dX.copy_(part[: self.block])

    def check_timeouts(self):
        if self.exchange == "p2p" and int(self.timed_out.item()) != 0:
            raise RuntimeError("a peer flag wait timed out: the exchange results are invalid")

    # -- measurement helpers used by bench.py -------------------------------------------------------------
    def kernel_times(self, iters=5):
        """Per-rank compute-only times (no exchange, no waits): forward over the row block; backward over all N nodes
        (peer blocks + own block) written locally."""
        out = {"fwd": [], "bwd": []}
        scratch = torch.empty((self.N, self.Dc), dtype=torch.float32, device=self.dev)
        for _ in range(iters):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            self._fwd()
            b.record()
            self._bwd_rows(scratch.data_ptr(), 0, self.N)
            c.record()
            torch.cuda.synchronize(self.dev)
            out["fwd"].append(a.elapsed_time(b))
            out["bwd"].append(b.elapsed_time(c))
        del scratch
        return out

    def exchange_times(self, iters=5):
        """Per-rank exchange-only times: the all-gather push alone and a plain peer copy of the partial volume."""
        if self.exchange != "p2p":
            return None
        out = {"all_gather_push": [], "partial_push": []}
        nbytes = self.n_rows * self.Dc * 4
        for _ in range(iters):
            dist.barrier()
            torch.cuda.synchronize(self.dev)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            main = torch.cuda.current_stream(self.dev)
            a.record()
            self._push_x(main)
            b.record()
            for i, q in enumerate(self.peers):          # same bytes as the fused partial stores, as a DMA
                qlo, qhi, _ = row_block(q, self.P_r, self.N)
                self._lib.check(self.lib.iddgcn_copy_async(self.peer_stage_dst[i], self.dX.data_ptr(),
                                                           min(nbytes, (qhi - qlo) * self.Dc * 4), main.cuda_stream))
            c.record()
            torch.cuda.synchronize(self.dev)
            out["all_gather_push"].append(a.elapsed_time(b))
            out["partial_push"].append(b.elapsed_time(c))
        dist.barrier()
        return out

    # -- end to end with host buffers -----------------------------------------------------------------------
    def step_host(self, Xh, dSh, Sh, dXh, chunks: int = 4, overlap: bool = True):
        """One step with pinned HOST blocks in and out: X_local/dS_local up, S_local/dX_local down.

        Three streams: uploads, compute, downloads.  X (the smaller input) goes up first so that the all-gather and
        the forward start early; S flows back in row chunks while dS is still on its way up; the backward's own block is
        cut into node chunks so that dX starts flowing back before the kernel has finished.
        overlap=False holds every download until both uploads have landed (the kernels still run as data arrives):
        on hosts where several GPUs share one PCIe uplink, simultaneous traffic in both directions from two GPUs is
        slower than the two directions one after the other (profiles/r02_host_step_timeline_c4_8gpu_2x4.jsonl)."""
        lib, chk = self.lib, self._lib.check
        main = torch.cuda.current_stream(self.dev)
        tl = getattr(self, "timeline", None)           # debug: list collecting (label, event) pairs of this step

        def mark(label, stream):
            if tl is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                tl.append((label, ev))

        if not hasattr(self, "s_in"):
            self.s_in = torch.cuda.Stream(device=self.dev)
            self.s_out = torch.cuda.Stream(device=self.dev)
        s_in, s_out = self.s_in, self.s_out
        p2p = self.exchange == "p2p"
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        xv = self.x_local_view()
        mark("start", s_in)
        chk(lib.iddgcn_copy_async(xv.data_ptr(), Xh.data_ptr(), xv.numel() * 4, s_in.cuda_stream))
        mark("X up", s_in)
        ev_x = s_in.record_event()
        if p2p:
            # the NVLink push goes on its own stream: s_in carries nothing but host->device copies and s_out nothing but
            # device->host copies (with the push queued between the two uploads the downloads were starved while dS
            # was on its way up: profiles/r02_host_step_timeline_1gpu_2gpu.jsonl)
            self.epoch += 1
            e = self.epoch
            side = self.s_side
            side.wait_event(ev_x)
            self._wait(K_DONE, e - 1, side)            # peers no longer read the X block I pushed in the previous step
            self._push_x(side)
            self._signal(K_AG, side)
        chk(lib.iddgcn_copy_async(self.dS.data_ptr(), dSh.data_ptr(), self.dS.numel() * 4, s_in.cuda_stream))
        mark("dS up", s_in)
        ev_ds = s_in.record_event()
        if not overlap:
            s_out.wait_event(ev_ds)
        main.wait_event(ev_x)
        if p2p:
            self._wait(K_AG, e)
        elif self.exchange == "nccl":
            dist.all_gather_into_tensor(self.X_full, self.X_full[self.pr * self.block:(self.pr + 1) * self.block],
                                        group=self.col_group)
        total = self.R * self.n_rows
        row_bytes = self.Dc * 4
        for q in range(chunks):
            lo, hi = total * q // chunks, total * (q + 1) // chunks
            self._fwd(lo, hi)
            mark(f"fwd chunk {q}", main)
            s_out.wait_event(main.record_event())
            chk(lib.iddgcn_copy_async(Sh.data_ptr() + lo * row_bytes, self.S.data_ptr() + lo * row_bytes,
                                      (hi - lo) * row_bytes, s_out.cuda_stream))
            mark(f"S down {q}", s_out)
        main.wait_event(ev_ds)
        if p2p:
            self._wait(K_DONE, e - 1)                  # peers are done with the stage buffers of the previous step
            self._push_partials()
            if self.part_local is not None:
                self._dma_partials(main)
            self._signal(K_PART)
            self._wait(K_PART, e)
        elif self.exchange == "nccl":
            self._bwd_rows(self.dX_partial.data_ptr(), 0, self.N)
            dist.reduce_scatter_tensor(self.dX, self.dX_partial, group=self.col_group)
        for q in range(chunks):
            a, b = self.n_rows * q // chunks, self.n_rows * (q + 1) // chunks
            if self.exchange != "nccl":
                self._own_partial(a, b)
            mark(f"bwd chunk {q}", main)
            s_out.wait_event(main.record_event())
            chk(lib.iddgcn_copy_async(dXh.data_ptr() + a * row_bytes, self.dX.data_ptr() + a * row_bytes,
                                      (b - a) * row_bytes, s_out.cuda_stream))
            mark(f"dX down {q}", s_out)
        if p2p:
            self._signal(K_DONE)
            main.wait_stream(self.s_side)
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    def e2e(self, steps, nnz_global):
        """bench.py's e2e leg at N > 1: every rank moves its own blocks over its own PCIe link; max over ranks."""
        Xh = self.x_local_view().cpu().pin_memory()
        dSh = self.dS.cpu().pin_memory()
        Sh = torch.empty(self.S.shape).pin_memory()
        dXh = torch.empty((self.n_rows, self.Dc)).pin_memory()
        k = max(3, min(steps, 5))
        # raw PCIe rate of this rank with every rank copying at once (what bounds the step)
        dist.barrier()
        torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        self.dS.copy_(dSh, non_blocking=True)
        torch.cuda.synchronize(self.dev)
        h2d = dSh.numel() * 4 / (time.perf_counter() - t0) / 1e9
        # copy schedule: both are tried (1 warm-up + 2 timed steps each, max over ranks) and the faster one is timed
        probe = {}
        for ov in (True, False):
            self.step_host(Xh, dSh, Sh, dXh, overlap=ov)
            torch.cuda.synchronize(self.dev)
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                self.step_host(Xh, dSh, Sh, dXh, overlap=ov)
            torch.cuda.synchronize(self.dev)
            t = torch.tensor([(time.perf_counter() - t0) / 2], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            probe[ov] = float(t[0])
        overlap = probe[True] <= probe[False]
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            self.step_host(Xh, dSh, Sh, dXh, overlap=overlap)
        torch.cuda.synchronize(self.dev)
        dt = torch.tensor([(time.perf_counter() - t0) / k, h2d, -h2d], dtype=torch.float64, device=self.dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        step_s, h2d_max, h2d_min = float(dt[0]), float(dt[1]), -float(dt[2])
        up = (Xh.numel() + dSh.numel()) * 4
        down = (Sh.numel() + dXh.numel()) * 4
        res = {"value": nnz_global / step_s, "unit": "edges/s", "h2d_bytes_per_step": up * self.world,
               "d2h_bytes_per_step": down * self.world, "ms_per_step": step_s * 1e3, "steps": k,
               "per_rank_pcie_GBps": {"up+down_in_step": (up + down) / step_s / 1e9,
                                      "h2d_alone_all_ranks_at_once_min": h2d_min,
                                      "h2d_alone_all_ranks_at_once_max": h2d_max},
               "copy_schedule": ("uploads and downloads overlapped" if overlap else
                                 "downloads held until the uploads have landed"),
               "copy_schedule_probe_ms": {"overlapped": probe[True] * 1e3, "uploads_then_downloads": probe[False] * 1e3},
               "api": "GridSpMM.step_host: pinned host blocks in/out on every rank (3-stream pipeline)"}
        del Xh, dSh, Sh, dXh
        return res
