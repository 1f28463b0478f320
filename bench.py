#!/usr/bin/env python3
"""bench.py — relational SpMM fwd+bwd throughput (edges/s) on B200, with roofline, CPU baseline and in-run parity.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|small|c3zipf]
                    [--grid RxC] [--exchange p2p|nccl]

A "step" is one pass of the hot path over one synthetic graph: S_r = A_r·X for every relation
(iddgcn_spmm_fwd) followed by dX = Σ_r A_rᵀ·dS_r (iddgcn_spmm_bwd_x).  edges = Σ_r nnz_r (deduplicated).
Workloads (BASELINE.json configs[2]/[3]; graph recipe: per relation r, np.random.default_rng(r)
.integers(0, N, per_rel) for heads then tails — the same function feeds both arms):
  c4    N=20M nodes, 2 x 100M raw edges, D=128 — BASELINE configs[3], the default at every --gpus N (strong
        scaling: the graph is fixed, the ranks form a P_r x P_c grid of destination-row blocks x feature-column
        groups, see iddgcn_b200/dist.py; it fits one B200, so N=1 runs the same graph).
  c3    N=1M nodes, 2 x 10M raw edges, D=128 — BASELINE configs[2]; with --gpus P the graph is P x c3 (weak scaling).
  small N=100k, 2 x 1M edges (smoke-sized).   c3zipf: c3 with Zipf(1.2) destination degrees.
Timing: CUDA events on the launching stream, W warm-up steps then K timed steps bracketed by barrier + synchronize;
max over ranks.  Inputs are far larger than the 126 MB L2, so no explicit flush is needed between iterations.
After the timed region every rank checks sampled rows of S and dX against oracle/spmm_ref.c evaluated on the RAW edge
list (independent of the device graph build and of the exchange) and the line carries "parity_check".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c4": dict(N=20_000_000, R=2, per_rel=100_000_000, D=128, scaling="strong"),
    "c3": dict(N=1_000_000, R=2, per_rel=10_000_000, D=128, scaling="weak"),
    "small": dict(N=100_000, R=2, per_rel=1_000_000, D=128, scaling="weak"),
    # skewed variant of c3: destination degrees ~ Zipf(s=1.2), same raw edge count
    "c3zipf": dict(N=1_000_000, R=2, per_rel=10_000_000, D=128, scaling="weak", zipf=1.2),
}
LABELS = {
    "c4": "synthetic 20M-node / 200M-edge 2-relation graph, D=128 (BASELINE configs[3])",
    "c3": "synthetic 1M-node / 20M-edge 2-relation directed graph, D=128 (BASELINE configs[2]); xP at --gpus P (weak)",
    "small": "synthetic 100k-node / 2M-edge 2-relation graph, D=128 (smoke size)",
    "c3zipf": "config-3 shape with Zipf(1.2) destination degrees (skewed variant, SURVEY 8d)",
}
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def algorithmic_bytes(nnz, n_rows, n_cols, R, D):
    """SURVEY §8(d) gather model: every edge fetches its 4D-byte source row and a 4-byte index."""
    fwd = nnz * (4 * D + 4) + R * n_rows * 4 * D + R * (n_rows + 1) * 4
    bwd = nnz * (4 * D + 4) + n_cols * 4 * D + R * (n_cols + 1) * 4
    return fwd, bwd


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic inputs shared by both arms
# ------------------------------------------------------------------------------------------------
def raw_edges(N, per_rel, r, zipf=None):
    """Relation r of the SURVEY §8d recipe: per_rel (head, tail) pairs, uniform (or Zipf-skewed heads), int64."""
    g = np.random.default_rng(r)
    if zipf:
        w = np.arange(1, N + 1, dtype=np.float64) ** -float(zipf)
        cdf = np.cumsum(w) / w.sum()
        scatter = np.random.default_rng(12345).permutation(N)
        h = scatter[np.minimum(np.searchsorted(cdf, g.random(per_rel)), N - 1)]
    else:
        h = g.integers(0, N, per_rel)
    t = g.integers(0, N, per_rel)
    return h, t


# Counter-based feature generator: element (row, col) of X or dS is a hash of its GLOBAL position, so any rank — and
# the host-side parity check — can produce any row without communication.  splitmix64 finaliser; torch (signed
# int64 with masked shifts) and NumPy (uint64) versions give identical bits.
_GOLD, _M1, _M2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
_STREAM = {"X": 0x1234 << 40, "dS": 0x4321 << 40}


def _s64(x):
    return x - (1 << 64) if x >= (1 << 63) else x


def _finish_np(u24, kind):
    u = u24.astype(np.float32) * np.float32(2.0 ** -24)
    return u if kind == "X" else (u - np.float32(0.5)) * np.float32(3.4641016)


def gen_rows_np(kind, row_ids, col_lo, Dc, D_total):
    """NumPy: rows `row_ids` (global flattened ids), columns [col_lo, col_lo+Dc) -> float32 [len, Dc]."""
    idx = (np.asarray(row_ids, np.uint64)[:, None] * np.uint64(D_total)
           + (np.uint64(col_lo) + np.arange(Dc, dtype=np.uint64))[None, :] + np.uint64(_STREAM[kind]))
    z = idx + np.uint64(_GOLD)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    z = z ^ (z >> np.uint64(31))
    return _finish_np(z >> np.uint64(40), kind)


def gen_rows_torch(torch, kind, row_ids, col_lo, Dc, D_total):
    """torch (any device): same values as gen_rows_np."""
    cols = torch.arange(col_lo, col_lo + Dc, dtype=torch.int64, device=row_ids.device)
    z = row_ids.to(torch.int64)[:, None] * D_total + cols[None, :] + (_STREAM[kind] + _s64(_GOLD))
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(_M1)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(_M2)
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    u = ((z >> 40) & 0xFFFFFF).to(torch.float32) * (2.0 ** -24)
    return u if kind == "X" else (u - 0.5) * 3.4641016


def fill_rows(torch, out, kind, first_row, col_lo, D_total, chunk_elems=1 << 25):
    """out [n, Dc] (device) <- generator rows first_row .. first_row+n."""
    n, Dc = out.shape
    step = max(1, chunk_elems // Dc)
    for a in range(0, n, step):
        b = min(n, a + step)
        ids = torch.arange(first_row + a, first_row + b, dtype=torch.int64, device=out.device)
        out[a:b].copy_(gen_rows_torch(torch, kind, ids, col_lo, Dc, D_total))


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle; the one place bench.py executes oracle/ besides the parity check)
# ------------------------------------------------------------------------------------------------
def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the
    OpenMP baseline to one thread)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_mem_short(need_bytes):
    """True when pinning need_bytes of host memory would leave less than 25 % + 8 GB of headroom."""
    try:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) * 1024 for l in f if l.startswith("MemAvailable"))
        limit = avail
        for p in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
            try:
                v = open(p).read().strip()
                if v.isdigit():
                    limit = min(limit, int(v))
            except OSError:
                pass
        host_mem_short.last = f"need {need_bytes / 1e9:.1f} GB, available {limit / 1e9:.1f} GB"
        return limit < 1.25 * need_bytes + (8 << 30)
    except Exception:
        host_mem_short.last = "unknown"
        return False


host_mem_short.last = ""


def cpu_baseline_run(struct, N, n_rows, R, D, X, dS, passes, threads=0):
    from oracle import cref
    if threads <= 0:
        threads = host_threads()
    S = np.empty((R * n_rows, D), np.float32)
    dX = np.empty((N, D), np.float32)
    times = []
    for _ in range(passes):
        t0 = time.perf_counter()
        cref.spmm_csr(struct["csr_rowptr"], struct["csr_col"], None, X, out=S, threads=threads)
        cref.spmm_csc_bwd(struct["csc_rowptr"], struct["csc_row"], struct["csc_perm"], None, R, N, n_rows, dS, out=dX,
                          threads=threads)
        times.append(time.perf_counter() - t0)
    return times, threads


def cpu_sample_inputs(wl):
    """Host-side inputs of the CPU arm: the SAME recipe (raw_edges + counter-hash features).  c3/small run the FULL
    workload (the very graph the GPU arm builds); c4 (200M edges) is bounded to the recipe at 1/8 scale."""
    frac = 8 if wl["per_rel"] > 20_000_000 else 1
    N, R, D = wl["N"] // frac, wl["R"], wl["D"]
    per_rel = wl["per_rel"] // frac
    rows_all, cols_all, perm_all, crow_all = [], [], [], []
    rowptr = np.zeros(R * N + 1, np.int64)
    cptr = np.zeros(R * N + 1, np.int64)
    off = 0
    for r in range(R):
        h, t = raw_edges(N, per_rel, r, wl.get("zipf"))
        key = np.unique(h * N + t)                      # dedup + (head, tail) order = CSR order
        rows, cols = key // N, key % N
        rowptr[r * N + 1:(r + 1) * N + 1] = off + np.cumsum(np.bincount(rows, minlength=N))
        order = np.argsort(cols * N + rows, kind="stable")      # (tail, head) order = CSC order
        cptr[r * N + 1:(r + 1) * N + 1] = off + np.cumsum(np.bincount(cols, minlength=N))
        rows_all.append(rows); cols_all.append(cols); perm_all.append(order + off); crow_all.append(rows[order])
        off += len(key)
    nnz = off
    struct = {"csr_rowptr": rowptr.astype(np.int32), "csr_col": np.concatenate(cols_all).astype(np.int32),
              "csc_rowptr": cptr.astype(np.int32), "csc_row": np.concatenate(crow_all).astype(np.int32),
              "csc_perm": np.concatenate(perm_all).astype(np.int32)}
    X = np.empty((N, D), np.float32)
    dS = np.empty((R * N, D), np.float32)
    step = 1 << 18
    for a in range(0, N, step):
        X[a:a + step] = gen_rows_np("X", np.arange(a, min(N, a + step)), 0, D, D)
    for a in range(0, R * N, step):
        dS[a:a + step] = gen_rows_np("dS", np.arange(a, min(R * N, a + step)), 0, D, D)
    scale = "full workload, same graph as the GPU arm" if frac == 1 else f"same recipe at 1/{frac} scale"
    sample = f"{wl['name']} ({scale}): {N} nodes, {nnz} edges, R={R}, D={D}, one fwd+bwd pass per step"
    return struct, N, nnz, X, dS, sample


def run_reference(args, wl):
    """--impl reference: the CPU restatement (oracle/spmm_ref.c, OpenMP over all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    struct, N, nnz, X, dS, sample = cpu_sample_inputs(wl)
    R, D = wl["R"], wl["D"]
    cpu_baseline_run(struct, N, N, R, D, X, dS, max(args.warmup, 1))
    times, cores = cpu_baseline_run(struct, N, N, R, D, X, dS, args.steps)
    total = sum(times)
    value = nnz * args.steps / total
    line = {
        "impl": "reference", "metric": "relational_spmm_fwd_bwd_edges_per_sec", "value": value, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["label"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# in-run parity check against the C oracle
# ------------------------------------------------------------------------------------------------
def parity_check(torch, eng, raw, N_total, R, D, n_sample, seed):
    """Sampled rows of this rank's S and dX against oracle/spmm_ref.c on the RAW edge list.

    Nothing the device computed enters the reference: the edges come from the host-side recipe (so the device graph
    build, dedup included, is under test), the input rows from the counter-hash generator (so the all-gather and the
    fused reduce-scatter are under test).  Returns per-rank maxima; bench reduces them over ranks."""
    from oracle import cref
    dev = eng.dev
    Dc, col_lo, lo, n_rows = eng.Dc, eng.pc * eng.Dc, eng.lo, eng.n_rows
    rng = np.random.default_rng(seed)
    m = min(n_sample, n_rows)
    if m <= 0:                                          # a rank that owns no rows (empty trailing block)
        return {"rows": 0, "edges": 0, "max_abs_err": 0.0, "max_rel_err": 0.0, "fwd_rel_err": 0.0, "bwd_rel_err": 0.0,
                "vs_tf_cpu_order_fp32": 0.0, "max_row_degree": 0}
    picked = np.sort(rng.choice(n_rows, m, replace=False)) if m < n_rows else np.arange(n_rows)
    gid = picked + lo                                   # global ids: destination rows for S, source nodes for dX
    lut = np.full(N_total, -1, np.int32)
    lut[gid] = np.arange(m, dtype=np.int32)
    threads = host_threads()

    def gen(kind, ids):                                 # generator rows on the device, then to the host
        out = np.empty((len(ids), Dc), np.float32)
        step = 1 << 20
        for a in range(0, len(ids), step):
            t = torch.from_numpy(np.ascontiguousarray(ids[a:a + step])).to(dev)
            out[a:a + step] = gen_rows_torch(torch, kind, t, col_lo, Dc, D).cpu().numpy()
        return out

    # --- forward: S[r, row] = Σ_{unique (row -> t) in relation r} X[t]
    ptr = np.zeros(R * m + 1, np.int64)
    cols = []
    for r, (h, t) in enumerate(raw):
        li = lut[h]
        sel = li >= 0
        key = np.unique(li[sel].astype(np.int64) * N_total + t[sel])
        ptr[r * m + 1:(r + 1) * m + 1] = np.bincount(key // N_total, minlength=m)
        cols.append(key % N_total)
    ptr = np.cumsum(ptr)
    cols = np.concatenate(cols)
    ucols, inv = np.unique(cols, return_inverse=True)
    p32, i32, Xs = ptr.astype(np.int32), inv.astype(np.int32), gen("X", ucols)
    S_ref = cref.spmm_csr_acc64(p32, i32, None, Xs, threads=threads)         # accurate reference (double accumulators)
    S_tf = cref.spmm_csr(p32, i32, None, Xs, threads=threads)                # TF-CPU order, fp32 accumulators
    sel_dev = torch.from_numpy(picked).to(dev)
    S_dev = eng.S[:, sel_dev].reshape(R * m, Dc).cpu().numpy()
    e_fwd = float(np.abs(S_dev - S_ref).max())
    s_fwd = float(np.abs(S_ref).max())
    tf_fwd = float(np.abs(S_dev - S_tf).max()) / (s_fwd + 1e-30)
    max_deg = int(np.diff(ptr).max())
    del Xs, S_tf
    n_edges = int(len(cols))
    # --- backward: dX[k] = Σ_r Σ_{unique (h -> k) in relation r} dS[r, h]   (relation-major, ascending head)
    keys = []
    for r, (h, t) in enumerate(raw):
        li = lut[t]
        sel = li >= 0
        keys.append(np.unique(li[sel].astype(np.int64) * (R * N_total) + (r * N_total + h[sel])))
    key = np.sort(np.concatenate(keys))
    ptr = np.concatenate([[0], np.cumsum(np.bincount(key // (R * N_total), minlength=m))])
    src = key % (R * N_total)
    usrc, inv = np.unique(src, return_inverse=True)
    p32, i32, Gs = ptr.astype(np.int32), inv.astype(np.int32), gen("dS", usrc)
    dX_ref = cref.spmm_csr_acc64(p32, i32, None, Gs, threads=threads)
    dX_tf = cref.spmm_csr(p32, i32, None, Gs, threads=threads)
    dX_dev = eng.dX[sel_dev].cpu().numpy()
    e_bwd = float(np.abs(dX_dev - dX_ref).max())
    s_bwd = float(np.abs(dX_ref).max())
    tf_bwd = float(np.abs(dX_dev - dX_tf).max()) / (s_bwd + 1e-30)
    max_deg = max(max_deg, int(np.diff(ptr).max()))
    n_edges += int(len(src))
    return {"rows": int(m), "edges": n_edges, "max_abs_err": max(e_fwd, e_bwd),
            "max_rel_err": max(e_fwd / (s_fwd + 1e-30), e_bwd / (s_bwd + 1e-30)),
            "fwd_rel_err": e_fwd / (s_fwd + 1e-30), "bwd_rel_err": e_bwd / (s_bwd + 1e-30),
            "vs_tf_cpu_order_fp32": max(tf_fwd, tf_bwd), "max_row_degree": max_deg}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from iddgcn_b200 import _lib, ops
    from iddgcn_b200.dist import GridSpMM, parse_grid, grid_coords, row_block
    from iddgcn_b200.graph import Graph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R, D = wl["R"], wl["D"]
    if wl["scaling"] == "weak":
        N_total, per_rel_total = wl["N"] * world, wl["per_rel"] * world
    else:
        N_total, per_rel_total = wl["N"], wl["per_rel"]
    grid = parse_grid(args.grid, world, D)
    pr, pc = grid_coords(rank, grid)
    row_lo, row_hi, block = row_block(pr, grid[0], N_total)
    n_rows = row_hi - row_lo
    Dc = D // grid[1]

    # graph: the raw edge list of the recipe (host), this rank's row block built on the device
    t0 = time.perf_counter()
    raw = [raw_edges(N_total, per_rel_total, r, wl.get("zipf")) for r in range(R)]
    gen_s = time.perf_counter() - t0
    parts = []
    for r, (h, t) in enumerate(raw):
        if grid[0] > 1:
            keep = (h >= row_lo) & (h < row_hi)
            h, t = h[keep], t[keep]
        tri = torch.empty((len(h), 3), dtype=torch.int64, device=dev)
        tri[:, 0].copy_(torch.from_numpy(h))
        tri[:, 1].fill_(r)
        tri[:, 2].copy_(torch.from_numpy(t))
        parts.append(tri)
    tri = torch.cat(parts, 0) if len(parts) > 1 else parts[0]
    del parts
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = Graph(tri, N_total, R, row_block=(row_lo, row_hi))
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    del tri
    torch.cuda.empty_cache()
    nnz_local = g.nnz

    eng = GridSpMM(g, D, world, rank, grid, exchange=args.exchange, partials=args.partials, order=args.order)
    fill_rows(torch, eng.x_local_view(), "X", row_lo, pc * Dc, D)
    for r in range(R):
        fill_rows(torch, eng.dS[r], "dS", r * N_total + row_lo, pc * Dc, D)
    # the two generators must agree bit for bit (the parity check relies on it)
    probe = np.array([0, 1, N_total - 1, 12345 % N_total], np.int64)
    for kind in ("X", "dS"):
        a = gen_rows_torch(torch, kind, torch.from_numpy(probe).to(dev), pc * Dc, Dc, D).cpu().numpy()
        assert np.array_equal(a, gen_rows_np(kind, probe, pc * Dc, Dc, D)), "device / host generators disagree"
    if world > 1:
        nnz_t = torch.tensor([nnz_local if pc == 0 else 0], device=dev, dtype=torch.int64)
        dist.all_reduce(nnz_t)
        nnz_global = int(nnz_t.item())
    else:
        nnz_global = nnz_local
    step = eng.step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # A timed region of K steps can be shorter than nvidia-smi's sampling period, so the clock sampler first watches
    # ~0.5 s of the same step (untimed, extra warm-up) and keeps running through the timed region that follows.
    t_load = time.perf_counter()
    step()
    torch.cuda.synchronize()
    est = torch.tensor([time.perf_counter() - t_load], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)        # every rank must run the same number of steps
    n_load = int(min(300, max(5, 0.5 / max(float(est.item()), 1e-4))))
    for _ in range(n_load):
        step()
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    eng.check_timeouts()
    total_ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = nnz_global / (ms_per_step * 1e-3)

    # in-run parity: every rank checks its share of --check-rows sampled rows (all rows when the block is smaller)
    parity = None
    if args.check_rows > 0:
        t0 = time.perf_counter()
        pch = parity_check(torch, eng, raw, N_total, R, D, max(1, args.check_rows // world), seed=1000 + rank)
        if world > 1:
            v = torch.tensor([pch["max_abs_err"], pch["max_rel_err"], pch["fwd_rel_err"], pch["bwd_rel_err"],
                              pch["vs_tf_cpu_order_fp32"], float(pch["max_row_degree"])], dtype=torch.float64, device=dev)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            c = torch.tensor([pch["rows"], pch["edges"]], dtype=torch.int64, device=dev)
            dist.all_reduce(c)
            pch.update(max_abs_err=float(v[0]), max_rel_err=float(v[1]), fwd_rel_err=float(v[2]),
                       bwd_rel_err=float(v[3]), vs_tf_cpu_order_fp32=float(v[4]), max_row_degree=int(v[5]),
                       rows=int(c[0]), edges=int(c[1]))
        pch.update(tolerance=1e-5, ok=bool(pch["max_rel_err"] <= 1e-5), ranks=world,
                   reference="oracle/spmm_ref.c on the raw edge list (sampled S rows and dX rows per rank, every rank): "
                             "max_rel_err against its double-accumulator flavour (the accurate value), "
                             "vs_tf_cpu_order_fp32 against TF-CPU's sequential fp32 loop (0.0 = bit-equal; a sequential "
                             "fp32 sum over a hub row of 10^5..10^6 edges is itself 1e-5 off, the GPU splits such rows)",
                   norm="max |gpu - ref| / max |ref|", seconds=time.perf_counter() - t0)
        parity = pch
    del raw

    # per-kernel timing (events on the launching stream) for the roofline of the dominant kernel
    kt = eng.kernel_times(iters=max(args.steps, 5))
    comm = None
    if world > 1:
        comm = {"grid": f"{grid[0]}x{grid[1]}", "exchange": eng.exchange, "partials": eng.partials, "order": eng.order,
                "bytes_exchanged_per_rank_each_way": int((grid[0] - 1) * block * Dc * 4)}
        et = eng.exchange_times(iters=5) if eng.exchange == "p2p" else None
        if et:
            comm["all_gather_push_ms_alone"] = float(np.median(et["all_gather_push"]))
            comm["partial_push_dma_ms_alone"] = float(np.median(et["partial_push"]))
            b = comm["bytes_exchanged_per_rank_each_way"]
            comm["all_gather_push_GBps"] = b / comm["all_gather_push_ms_alone"] / 1e6
            # NVLink bound (B200_PROFILING.md: ~770 GB/s per direction per GPU): each rank sends its X block and its
            # partials to P_r-1 peers -> 2 x bytes out, and receives as much
            comm["nvlink_bound_ms"] = 2 * b / 770.0 / 1e6
        ph = eng.step_phases(iters=5)
        if ph:
            comm["phase_done_at_ms_rank0"] = {k: round(v, 3) for k, v in ph.items()}
        fwd_k, bwd_k = float(np.median(kt["fwd"])), float(np.median(kt["bwd"]))
        comm["kernels_ms"] = fwd_k + bwd_k
        comm["step_over_kernels"] = ms_per_step / (fwd_k + bwd_k)
        comm["step_minus_kernels_ms"] = ms_per_step - (fwd_k + bwd_k)
        if ph and eng.order == "ABF":
            a0 = ph.get("peers done with previous step", 0.0)
            a_ms = ph["A: peers' partials"] - a0
            b_ms = ph["B: own block + staged partials"] - ph["partials arrived"]
            f_ms = ph["F: forward"] - ph["X blocks arrived"]
            push_ms = ph["X block pushed (side)"] - ph["A: peers' partials"]
            b_ex = comm["bytes_exchanged_per_rank_each_way"]
            a_compute = bwd_k * (grid[0] - 1) / grid[0]
            comm["phases_ms"] = {"A: backward over the peers' node blocks, rows stored over NVLink": a_ms,
                                 "B: backward over the own block + staged partials (X push underneath)": b_ms,
                                 "X push on the copy engines (under B)": push_ms, "F: forward": f_ms}
            comm["limiter"] = (f"A is NVLink-bound: {b_ex / a_ms / 1e6:.0f} GB/s for {b_ex / 1e9:.2f} GB of partial rows "
                               f"({a_ms:.2f} ms; the same launches write locally in ~{a_compute:.2f} ms); B {b_ms:.2f} ms "
                               f"vs X push {push_ms:.2f} ms run together; F {f_ms:.2f} ms needs the pushed blocks, so the "
                               f"step is A + max(B, push) + F = {a_ms + max(b_ms, push_ms) + f_ms:.2f} ms of "
                               f"{ms_per_step:.2f} ms")
        else:
            comm["limiter"] = ("kernels (no exchange, or exchange hidden)" if ms_per_step < 1.15 * (fwd_k + bwd_k) else
                               "exchange / waits not fully hidden behind the kernels")
    sddmm = None
    if world == 1 and not args.no_sddmm:
        # separately reported explainer line: edge-value gradient dvals[e] = <dS[row_e], X[col_e]>
        dv = torch.empty(nnz_local, dtype=torch.float32, device=dev)
        Xv, dSv = eng.x_local_view(), eng.dS
        for _ in range(5):
            ops.spmm_bwd_vals(g, dSv, Xv, out=dv)
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.spmm_bwd_vals(g, dSv, Xv, out=dv); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        del dv, Xv, dSv
        sd_ms = float(np.median(ts))
        bv = nnz_local * (8 * D + 8 + 4)
        sddmm = {"kernel": "sddmm_stream_kernel (iddgcn_spmm_bwd_vals)", "ms": sd_ms, "ms_min": float(min(ts)),
                 "ms_max": float(max(ts)), "iters": 20, "bytes": bv, "GBps": bv / sd_ms / 1e6,
                 "edges_per_s": nnz_local / (sd_ms * 1e-3)}
    fwd_ms, bwd_ms = float(np.median(kt["fwd"])), float(np.median(kt["bwd"]))
    bf, bb = algorithmic_bytes(nnz_local, n_rows, N_total, R, Dc)
    peak, peak_src = hbm_peak()
    dom = "fwd" if fwd_ms >= bwd_ms else "bwd"
    dom_bytes, dom_ms = (bf, fwd_ms) if dom == "fwd" else (bb, bwd_ms)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    if world == 1 and not wl.get("dim_override"):
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get(wl["name"], {}).get(dom)
            if traffic is not None:
                traffic_src = "static: ncu --set full capture of this kernel on this workload, committed in profiles/"
        except Exception:
            pass
    kname = "gather_rows_kernel" if os.environ.get("IDDGCN_SPMM_V1") == "1" else "stream_rows_kernel"
    roofline = {"bound": "hbm", "kernel": f"{kname} ({'iddgcn_spmm_fwd' if dom == 'fwd' else 'iddgcn_spmm_bwd_x'})",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes": dom_bytes, "ms": dom_ms,
                "per_rank_shape": {"rows": n_rows, "source_nodes": N_total, "edges": nnz_local, "feature_cols": Dc},
                "fwd": {"ms": fwd_ms, "bytes": bf, "GBps": bf / fwd_ms / 1e6, "frac": bf / fwd_ms / 1e6 / peak},
                "bwd": {"ms": bwd_ms, "bytes": bb, "GBps": bb / bwd_ms / 1e6, "frac": bb / bwd_ms / 1e6 / peak},
                "step_frac": (bf + bb) / (ms_per_step * 1e-3) / 1e9 / peak}

    # end-to-end through the C ABI with HOST buffers (pinned): H2D + fwd + bwd + D2H every step
    e2e = None
    if args.no_e2e:
        e2e = {"value": None, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": True}
    elif world == 1 and host_mem_short(4 * (2 * N_total * D + 2 * R * n_rows * D)):
        e2e = {"value": None, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": True,
               "reason": "host RAM too small to pin the buffers of this workload: " + host_mem_short.last}
    elif world == 1:
        Xh = torch.empty((N_total, D), pin_memory=True); Xh.copy_(eng.x_local_view())
        dSh = torch.empty((R, n_rows, D), pin_memory=True); dSh.copy_(eng.dS)
        Sh = torch.empty((R, n_rows, D), pin_memory=True); dXh = torch.empty((N_total, D), pin_memory=True)
        del eng, step                                 # the host entry point owns its device buffers
        torch.cuda.empty_cache()
        k2 = max(3, min(args.steps, 5))
        for _ in range(2):
            ops.spmm_fwd_bwd_host(g, Xh, dSh, Sh, dXh)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(k2):
            ops.spmm_fwd_bwd_host(g, Xh, dSh, Sh, dXh)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k2
        e2e = {"value": nnz_global / dt, "unit": "edges/s", "h2d_bytes_per_step": Xh.numel() * 4 + dSh.numel() * 4,
               "d2h_bytes_per_step": Sh.numel() * 4 + dXh.numel() * 4, "ms_per_step": dt * 1e3, "steps": k2,
               "launches_per_step": (_lib.launch_count() - l0) / k2,
               "pcie_GBps_up+down": (Xh.numel() + dSh.numel() + Sh.numel() + dXh.numel()) * 4 / dt / 1e9,
               "api": "iddgcn_spmm_fwd_bwd_host (pinned host buffers in/out)"}
        del Xh, dSh, Sh, dXh
    else:
        e2e = eng.e2e(args.steps, nnz_global)
        eng.check_timeouts()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        struct, cN, cnnz, cX, cdS, sample = cpu_sample_inputs(wl)
        cpu_baseline_run(struct, cN, cN, R, D, cX, cdS, 1)
        times, cores = cpu_baseline_run(struct, cN, cN, R, D, cX, cdS, 2)
        cpu = {"value": cnnz * len(times) / sum(times), "unit": "edges/s", "cores": cores, "kind": "port",
               "sample": sample + ", 2 timed passes", "host_cpus": os.cpu_count()}
    if rank == 0:
        line = {
            "metric": "relational_spmm_fwd_bwd_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "nodes": N_total, "edges": nnz_global, "relations": R,
                       "feature_dim": D, "rows_per_gpu": n_rows, "feature_cols_per_gpu": Dc,
                       "partition": (f"{grid[0]} destination-row blocks x {grid[1]} feature-column groups; rank (pr,pc) "
                                     f"holds X[rows_pr, cols_pc], dS/S[:, rows_pr, cols_pc], dX[rows_pr, cols_pc]"),
                       "l2": "inputs larger than L2 (no flush needed)", "graph_build_s": build_s,
                       "edge_recipe_s": gen_s},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_check": parity,
        }
        if comm is not None:
            line["comm"] = comm
        if sddmm is not None:
            sddmm["frac"] = sddmm["GBps"] / peak
            line["sddmm"] = sddmm
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (A/B runs only)")
    ap.add_argument("--no-sddmm", action="store_true", help="skip the separately reported SDDMM line")
    ap.add_argument("--dim", type=int, default=0, help="override the feature width D of the workload (default 128)")
    ap.add_argument("--grid", default="auto", help="rank grid 'RxC': R destination-row blocks x C feature-column "
                                                    "groups (auto: 1x2, 1x4, 2x4 at 2/4/8 GPUs)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="row exchange inside a column group: fused peer-memory kernels or NCCL collectives")
    ap.add_argument("--partials", default="stores", choices=["stores", "dma"],
                    help="p2p exchange: partial dX rows leave through NVLink stores of the backward kernel (fused) or "
                         "through the copy engines")
    ap.add_argument("--order", default="ABF", choices=["AFB", "ABF"],
                    help="p2p exchange: kernel order inside a step (A peers' backward blocks, F forward, B own block)")
    ap.add_argument("--check-rows", type=int, default=1_000_000,
                    help="rows (over all ranks) of S and of dX compared with the CPU oracle after the timed region")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = dict(WORKLOADS[args.workload])
    wl["name"] = args.workload
    if args.dim:
        wl["D"] = args.dim
        wl["dim_override"] = True
    wl["label"] = LABELS[args.workload]
    if wl.get("dim_override"):
        wl["label"] += f" [feature dim overridden to {wl['D']}]"
    if args.impl == "reference":
        run_reference(args, wl)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun when called as plain `python bench.py --gpus N`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, wl)


if __name__ == "__main__":
    main()
