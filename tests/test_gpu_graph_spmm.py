"""GPU parity: device graph construction (bit-exact) and the relational SpMM family, through the C ABI.

Oracle = oracle/iddgcn_oracle.py (CPU restatement of utils1.get_adj_mats and of
tf.sparse.sparse_dense_matmul + gradient).  Integer structure must be bit-exact; fp32 kernels must be
within 1e-5 relative (north_star) of the fp64 oracle, measured against the magnitude of the result.
"""
import numpy as np
import pytest
import torch

from helpers import random_triples, rel_err
from oracle import iddgcn_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _oracle_struct(tri, N, R, lo=0, hi=None):
    return O.csr_csc_from_adj(O.get_adj_mats(tri, N, R), N, lo, hi)


@pytest.mark.parametrize("N,R,M,kw", [
    (23, 3, 90, {}),
    (845, 4, 40000, {"heavy_row": 7, "bad_rel": 11}),
    (1000, 5, 3000, {"empty_rel": 2}),
    (1, 1, 5, {}),
    (70000, 2, 500000, {"heavy_row": 69999}),
    (1 << 16, 4, 200000, {}),          # head/tail/rel fields exactly fill their bit widths
])
def test_graph_build_bit_exact(N, R, M, kw):
    from iddgcn_b200.graph import Graph
    tri = random_triples(N, R, M, seed=N + M, **kw)
    g = Graph(tri, N, R)
    ref = _oracle_struct(tri, N, R)
    got = g.export()
    assert g.nnz == ref["nnz"]
    adj = O.get_adj_mats(tri, N, R)
    assert g.nnz_per_rel == [0 if _is_placeholder(a) else a.nnz for a in adj]
    for k in ("csr_rowptr", "csr_col", "csr_row", "csc_rowptr", "csc_row", "csc_perm"):
        assert np.array_equal(got[k], ref[k]), k
    # device-side triples give the same structure
    g2 = Graph(torch.from_numpy(tri).cuda(), N, R)
    got2 = g2.export()
    for k in got:
        assert np.array_equal(got[k], got2[k]), k


def test_graph_build_empty_and_errors():
    from iddgcn_b200._lib import IddgcnError
    from iddgcn_b200.graph import Graph, get_adj_mats
    g = Graph(np.zeros((0, 3), np.int64), 10, 2)
    assert g.nnz == 0 and g.export()["csr_rowptr"].tolist() == [0] * 21
    with pytest.raises(IddgcnError) as ei:
        Graph(np.array([[0, 0, 10]], np.int64), 10, 2)
    assert ei.value.code == -4
    # reference placeholder for an empty relation: indices [[0,0,0]], values [0.] (utils1.py:427-429)
    adj = get_adj_mats(np.array([[1, 0, 2], [1, 0, 2], [3, 0, 1]]), 5, 2)
    ref = O.get_adj_mats(np.array([[1, 0, 2], [1, 0, 2], [3, 0, 1]]), 5, 2)
    for a, b in zip(adj, ref):
        assert np.array_equal(a.indices.cpu().numpy(), b.indices)
        assert np.array_equal(a.values.cpu().numpy(), b.values)
        assert a.dense_shape == b.dense_shape


def test_graph_row_block_partition():
    from iddgcn_b200.graph import Graph
    N, R = 1000, 3
    tri = random_triples(N, R, 20000, seed=5)
    for lo, hi in [(0, 250), (250, 777), (777, 1000)]:
        g = Graph(tri, N, R, row_block=(lo, hi))
        ref = _oracle_struct(tri, N, R, lo, hi)
        got = g.export()
        assert g.n_rows == hi - lo and g.nnz == ref["nnz"]
        for k in ("csr_rowptr", "csr_col", "csr_row", "csc_rowptr", "csc_row", "csc_perm"):
            assert np.array_equal(got[k], ref[k]), (k, lo, hi)


def _is_placeholder(a):
    return a.nnz == 1 and a.values[0] == 0


def _vals_for(adj, mode, seed):
    """raw: float32 vector the kernels get (CSR edge order); eff: per-relation fp64 effective edge
    values for the oracle (the reference's placeholder entry keeps its single 0.0)."""
    if mode == 0:
        return None, None
    g = np.random.default_rng(seed)
    raw = [g.standard_normal(0 if _is_placeholder(a) else a.nnz).astype(np.float32) for a in adj]
    eff = []
    for r_, a in zip(raw, adj):
        e = r_.astype(np.float64)
        if mode == 2:
            e = O.sigmoid(e)
        eff.append(np.zeros(1) if _is_placeholder(a) else e)
    return np.concatenate(raw), eff


@pytest.mark.parametrize("kernels", ["v1", "v2"])
@pytest.mark.parametrize("D", [4, 8, 16, 32, 64, 100, 128, 256, 320])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_spmm_family_parity(D, mode, kernels, monkeypatch):
    """Both kernel generations: v1 = one row per lane slice (small graphs, D > 128), v2 = shared-memory index slabs +
    flat edge stream (what the bench sizes run); IDDGCN_SPMM_V1 forces the choice regardless of the row count."""
    monkeypatch.setenv("IDDGCN_SPMM_V1", "1" if kernels == "v1" else "0")
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R, M = 777, 3, 9000
    tri = random_triples(N, R, M, seed=D + mode, heavy_row=3, empty_rel=None if D != 64 else 1)
    adj = O.get_adj_mats(tri, N, R)
    g = Graph(tri, N, R)
    rng = np.random.default_rng(D)
    X = rng.random((N, D))
    dS = rng.standard_normal((R, N, D))
    raw, eff = _vals_for(adj, mode, seed=3)
    Xd = torch.from_numpy(X.astype(np.float32)).cuda()
    dSd = torch.from_numpy(dS.astype(np.float32)).cuda()
    vd = torch.from_numpy(raw).cuda() if raw is not None else None
    X32, dS32 = X.astype(np.float32).astype(np.float64), dS.astype(np.float32).astype(np.float64)
    eff64 = eff
    S_ref = O.spmm_fwd(adj, X32, eff64)
    S = ops.spmm_fwd(g, Xd, vd, mode).cpu().numpy()
    assert rel_err(S, S_ref) < TOL
    dX_ref = O.spmm_bwd_x(adj, dS32, eff64)
    dX = ops.spmm_bwd_x(g, dSd, vd, mode).cpu().numpy()
    assert rel_err(dX, dX_ref) < TOL
    # accumulate flag
    base = torch.ones_like(Xd)
    dX2 = ops.spmm_bwd_x(g, dSd, vd, mode, out=base.clone(), accumulate=True).cpu().numpy()
    assert rel_err(dX2, dX_ref + 1.0) < TOL
    # SDDMM (edge-value gradient); scale = Σ|terms| bound, since the dot products cancel
    dv_ref = O.spmm_bwd_vals(adj, dS32, X32)
    dv_ref = np.concatenate([d for d, a in zip(dv_ref, adj) if not _is_placeholder(a)])
    if mode == 2:
        s = np.concatenate([e for e, a in zip(eff64, adj) if not _is_placeholder(a)])
        dv_ref = dv_ref * s * (1 - s)
    dv = ops.spmm_bwd_vals(g, dSd, Xd, vd, mode).cpu().numpy()
    assert dv.shape == dv_ref.shape
    assert rel_err(dv, dv_ref, scale=np.abs(dS32).max() * np.abs(X32).max() * np.sqrt(D)) < TOL


def test_spmm_reproducible_and_dlpack_and_host():
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R, D = 5000, 2, 128
    tri = random_triples(N, R, 60000, seed=11)
    g = Graph(tri, N, R)
    X = torch.rand(N, D, device="cuda")
    dS = torch.randn(R, N, D, device="cuda")
    S1 = ops.spmm_fwd(g, X)
    S2 = ops.spmm_fwd(g, X)
    assert torch.equal(S1, S2)                      # atomics-free => bit-reproducible
    S3 = torch.empty_like(S1)
    ops.spmm_fwd_dlpack(g, X, S3)
    assert torch.equal(S1, S3)
    dX1 = ops.spmm_bwd_x(g, dS)
    dX3 = torch.empty_like(dX1)
    ops.spmm_bwd_x_dlpack(g, dS, dX3)
    assert torch.equal(dX1, dX3)
    Xh, dSh = X.cpu().pin_memory(), dS.cpu().pin_memory()
    Sh = torch.empty(R, N, D).pin_memory()
    dXh = torch.empty(N, D).pin_memory()
    ops.spmm_fwd_bwd_host(g, Xh, dSh, Sh, dXh)
    assert torch.equal(Sh, S1.cpu()) and torch.equal(dXh, dX1.cpu())
    # DLPack path refuses CPU tensors loudly (no CPU fallback)
    from iddgcn_b200._lib import IddgcnError
    with pytest.raises(IddgcnError):
        ops.spmm_fwd_dlpack(g, X.cpu(), S3)


@pytest.mark.parametrize("panels", ["1", "2", "4"])
@pytest.mark.parametrize("D", [128, 64, 100, 8])
def test_host_path_column_panels_are_bit_identical(D, panels, monkeypatch):
    """iddgcn_spmm_fwd_bwd_host moves the buffers as column panels (strided 2-D copies) to keep both PCIe directions
    busy; whatever the panel count — including widths it cannot split evenly and a hub row — the host results equal the
    device-resident ones bit for bit, and unrequested outputs (NULL S / dX) are skipped."""
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    monkeypatch.setenv("IDDGCN_HOST_PANELS", panels)
    N, R = 3000, 3
    tri = random_triples(N, R, 40000, seed=5, heavy_row=2)
    g = Graph(tri, N, R)
    X = torch.rand(N, D, device="cuda")
    dS = torch.randn(R, N, D, device="cuda")
    S_ref, dX_ref = ops.spmm_fwd(g, X), ops.spmm_bwd_x(g, dS)
    Xh, dSh = X.cpu().pin_memory(), dS.cpu().pin_memory()
    Sh = torch.full((R, N, D), float("nan")).pin_memory()
    dXh = torch.full((N, D), float("nan")).pin_memory()
    ops.spmm_fwd_bwd_host(g, Xh, dSh, Sh, dXh)
    assert torch.equal(Sh, S_ref.cpu()) and torch.equal(dXh, dX_ref.cpu())
    Sh2 = torch.full((R, N, D), float("nan"))                                    # pageable memory works too
    ops.spmm_fwd_bwd_host(g, X.cpu(), dS.cpu(), Sh2, None)
    assert torch.equal(Sh2, S_ref.cpu())


def test_spmm_autograd_matches_oracle():
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R, D = 300, 4, 64
    tri = random_triples(N, R, 4000, seed=2)
    adj = O.get_adj_mats(tri, N, R)
    g = Graph(tri, N, R)
    rng = np.random.default_rng(0)
    X = torch.tensor(rng.random((N, D)), dtype=torch.float32, device="cuda", requires_grad=True)
    v = torch.tensor(rng.uniform(0.5, 1.5, g.nnz), dtype=torch.float32, device="cuda", requires_grad=True)
    w = torch.tensor(rng.standard_normal((R, N, D)), dtype=torch.float32, device="cuda")
    S = ops.relational_spmm(g, X, v, 1)
    (S * w).sum().backward()
    vals = np.split(v.detach().cpu().numpy().astype(np.float64), np.cumsum(g.nnz_per_rel)[:-1])
    dX_ref = O.spmm_bwd_x(adj, w.cpu().numpy().astype(np.float64), vals)
    dv_ref = np.concatenate(O.spmm_bwd_vals(adj, w.cpu().numpy().astype(np.float64),
                                            X.detach().cpu().numpy().astype(np.float64)))
    assert rel_err(X.grad.cpu().numpy(), dX_ref) < TOL
    assert rel_err(v.grad.cpu().numpy(), dv_ref, scale=np.abs(dv_ref).max() + 10) < TOL


def test_large_graph_properties():
    """Full bench-size (C3) checks through size-independent properties: linearity, adjoint identity
    <A·X, dS> = <X, Aᵀ·dS>, row-sum identity against the CSR row pointer, sortedness of the structure."""
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R, D, per_rel = 1_000_000, 2, 128, 10_000_000
    torch.manual_seed(0)
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randint(0, N, (R * per_rel,), device="cuda", generator=gen)
    t = torch.randint(0, N, (R * per_rel,), device="cuda", generator=gen)
    r = torch.arange(R, device="cuda").repeat_interleave(per_rel)
    g = Graph(torch.stack([h, r, t], 1), N, R)
    d = g.device_arrays()
    # structure: keys strictly increasing inside the global (rel, head, tail) order, CSC is a permutation
    seg = torch.repeat_interleave(torch.arange(R * N, device="cuda"),
                                  (d["csr_rowptr"][1:] - d["csr_rowptr"][:-1]).long())
    key = seg * N + d["csr_col"].long()
    assert bool((key[1:] > key[:-1]).all())
    perm = d["csc_perm"].long()
    assert bool((torch.sort(perm).values == torch.arange(g.nnz, device="cuda")).all())
    segc = torch.repeat_interleave(torch.arange(R * N, device="cuda"),
                                   (d["csc_rowptr"][1:] - d["csc_rowptr"][:-1]).long())
    # CSC entry j is CSR edge perm[j]: same (rel, tail) segment and same head
    assert bool((segc % N == d["csr_col"].long()[perm]).all())
    assert bool((d["csc_row"].long() == d["csr_row"].long()[perm]).all())
    del seg, key, segc, perm
    X = torch.rand(N, D, device="cuda")
    ones = torch.ones(N, D, device="cuda")
    S1 = ops.spmm_fwd(g, ones)
    deg = (d["csr_rowptr"][1:] - d["csr_rowptr"][:-1]).float().view(R, N, 1)
    assert torch.equal(S1, deg.expand(R, N, D))      # A·1 = out-degree, exact in fp32
    del S1, ones
    S = ops.spmm_fwd(g, X)
    S2 = ops.spmm_fwd(g, 2 * X)
    assert torch.equal(S2, 2 * S)                    # scaling by 2 is exact
    del S2
    dS = torch.randn(R, N, D, device="cuda")
    dX = ops.spmm_bwd_x(g, dS)
    lhs = (S.double() * dS.double()).sum().item()
    rhs = (X.double() * dX.double()).sum().item()
    # both sides are sums of 2.6e8 random-sign terms computed from fp32 S / dX: the rounding noise of such a sum is
    # ~1e-7·sqrt(numel)·rms(S)·rms(dS) (~0.01 here) while a structural error (e.g. a transposition) is ~sqrt(numel)·rms
    noise = 1e-7 * (S.numel() ** 0.5) * S.double().pow(2).mean().sqrt().item() * dS.double().pow(2).mean().sqrt().item()
    assert abs(lhs - rhs) <= 50 * noise, (lhs, rhs, noise)
    # SDDMM at full size: Σ_e dvals[e] = <A·X, dS> (values all 1), and a sample of edges against direct dot products
    dv = ops.spmm_bwd_vals(g, dS, X)
    assert abs(dv.double().sum().item() - lhs) <= 50 * noise
    e = torch.randint(0, g.nnz, (100000,), device="cuda")
    rel = torch.bucketize(e, torch.from_numpy(g.rel_offsets).cuda()[1:], right=True)
    ref = (dS[rel, d["csr_row"].long()[e]].double() * X[d["csr_col"].long()[e]].double()).sum(1)
    assert (dv[e].double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("kernels", ["v1", "v2"])
@pytest.mark.parametrize("D,mode", [(128, 0), (64, 2), (256, 1)])
def test_hub_rows_take_the_split_pass(D, mode, kernels, monkeypatch):
    """Rows / columns with more than 1024 edges are cut into segments summed by separate warps (Zipf-skewed graphs):
    a hub destination row (forward, SDDMM), a hub source column (backward), one of them spanning several segments."""
    monkeypatch.setenv("IDDGCN_SPMM_V1", "1" if kernels == "v1" else "0")
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R = 30000, 2
    rng = np.random.default_rng(D + mode)
    base = random_triples(N, R, 50000, seed=D)
    hub_row = np.stack([np.full(20000, 5), np.zeros(20000, np.int64), rng.integers(0, N, 20000)], 1)
    hub_row2 = np.stack([np.full(6000, N - 1), np.ones(6000, np.int64), rng.integers(0, N, 6000)], 1)
    hub_col = np.stack([rng.integers(0, N, 15000), np.ones(15000, np.int64), np.full(15000, 7)], 1)
    tri = np.concatenate([base, hub_row, hub_row2, hub_col]).astype(np.int64)
    adj = O.get_adj_mats(tri, N, R)
    g = Graph(tri, N, R)
    ref = _oracle_struct(tri, N, R)
    got = g.export()
    for k in ("csr_rowptr", "csr_col", "csc_rowptr", "csc_row", "csc_perm"):
        assert np.array_equal(got[k], ref[k]), k
    deg = np.diff(ref["csr_rowptr"])
    assert (deg > 1024).sum() >= 2 and deg.max() > 3 * 1024
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    raw, eff = _vals_for(adj, mode, seed=1)
    Xd, dSd = torch.from_numpy(X).cuda(), torch.from_numpy(dS).cuda()
    vd = torch.from_numpy(raw).cuda() if raw is not None else None
    S = ops.spmm_fwd(g, Xd, vd, mode).cpu().numpy()
    assert rel_err(S, O.spmm_fwd(adj, X.astype(np.float64), eff)) < TOL
    dX = ops.spmm_bwd_x(g, dSd, vd, mode).cpu().numpy()
    assert rel_err(dX, O.spmm_bwd_x(adj, dS.astype(np.float64), eff)) < TOL
    dX2 = ops.spmm_bwd_x(g, dSd, vd, mode, out=torch.ones_like(Xd), accumulate=True).cpu().numpy()
    assert rel_err(dX2, O.spmm_bwd_x(adj, dS.astype(np.float64), eff) + 1.0) < TOL
    dv_ref = np.concatenate(O.spmm_bwd_vals(adj, dS.astype(np.float64), X.astype(np.float64)))
    if mode == 2:
        sg = np.concatenate(eff)
        dv_ref = dv_ref * sg * (1 - sg)
    dv = ops.spmm_bwd_vals(g, dSd, Xd, vd, mode).cpu().numpy()
    assert rel_err(dv, dv_ref, scale=np.abs(dS).max() * np.abs(X).max() * np.sqrt(D)) < TOL
    assert torch.equal(ops.spmm_fwd(g, Xd, vd, mode), torch.from_numpy(S).cuda())     # still bit-reproducible


@pytest.mark.parametrize("D", [16, 32, 128])
def test_row_ranges_addends_and_hub_rows_v2(D, monkeypatch):
    """The range entry points the pipelined / partitioned callers use (iddgcn_spmm_fwd_rows, iddgcn_spmm_bwd_x_rows with
    epilogue addends) on a graph with hub rows AND hub columns, v2 kernels forced: chunked calls reproduce the
    one-shot result bit for bit, addends are summed in order, and the v2 SDDMM skips / fills hub rows correctly."""
    monkeypatch.setenv("IDDGCN_SPMM_V1", "0")
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R = 6000, 2
    tri = random_triples(N, R, 50000, seed=D, heavy_row=11)
    hub_col = tri[:9000].copy()
    hub_col[:, 2] = 17                              # > 1024 in-edges of node 17: a hub on the backward side too
    empties = tri[(tri[:, 0] % 7) != 3]             # whole runs of empty rows between non-empty ones
    tri = np.concatenate([empties, hub_col], 0)
    adj = O.get_adj_mats(tri, N, R)
    g = Graph(tri, N, R)
    rng = np.random.default_rng(D)
    X = rng.random((N, D)).astype(np.float32)
    dS = rng.standard_normal((R, N, D)).astype(np.float32)
    Xd, dSd = torch.from_numpy(X).cuda(), torch.from_numpy(dS).cuda()
    S_full = ops.spmm_fwd(g, Xd)
    assert rel_err(S_full.cpu().numpy(), O.spmm_fwd(adj, X.astype(np.float64))) < TOL
    S_chunks = torch.full_like(S_full, -7.0)
    total = R * N
    for lo, hi in [(0, 1), (1, 4001), (4001, 4001), (4001, total)]:
        ops.spmm_fwd_rows(g, Xd, S_chunks, lo, hi)
    assert torch.equal(S_chunks, S_full)
    dX_full = ops.spmm_bwd_x(g, dSd)
    assert rel_err(dX_full.cpu().numpy(), O.spmm_bwd_x(adj, dS.astype(np.float64))) < TOL
    a1 = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).cuda()
    a2 = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).cuda()
    out = torch.full((N, D), -7.0, device="cuda")
    for lo, hi in [(0, 18), (18, 3000), (3000, N)]:
        ops.spmm_bwd_x_rows(g, dSd, out[lo:], lo, hi, addends=[a1[lo:], a2[lo:]])
    assert torch.equal(out, dX_full + (a1 + a2))        # addends are summed in order, then added to the gathered sum
    plain = torch.empty((N, D), device="cuda")
    ops.spmm_bwd_x_rows(g, dSd, plain, 0, N)
    assert torch.equal(plain, dX_full)
    # SDDMM with hub rows, three modes (ones / explicit share the kernel)
    for mode in (1, 2):
        raw, eff = _vals_for(adj, mode, seed=5)
        vd = torch.from_numpy(raw).cuda()
        dv_ref = np.concatenate(O.spmm_bwd_vals(adj, dS.astype(np.float64), X.astype(np.float64)))
        if mode == 2:
            s_ = np.concatenate(eff)
            dv_ref = dv_ref * s_ * (1 - s_)
        dv = ops.spmm_bwd_vals(g, dSd, Xd, vd, mode).cpu().numpy()
        assert rel_err(dv, dv_ref, scale=np.abs(dS).max() * np.abs(X).max() * np.sqrt(D)) < TOL


def test_empty_row_block_graph():
    """A trailing rank of a partition may own no rows (N not divisible by the number of blocks): the handle builds,
    reports nnz = 0, and the backward over its (empty) block writes zeros."""
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    tri = random_triples(50, 2, 400, seed=1)
    g = Graph(tri, 50, 2, row_block=(50, 50))
    assert g.n_rows == 0 and g.nnz == 0
    dX = ops.spmm_bwd_x(g, torch.zeros((2, 0, 8), device="cuda"))
    assert dX.shape == (50, 8) and float(dX.abs().max()) == 0.0


@pytest.mark.parametrize("D", [8, 16, 24, 32, 64, 128])
def test_addends_without_hub_rows_all_widths(D, monkeypatch):
    """Epilogue addends on a graph WITHOUT hub rows (the instantiation the partitioned backward runs at bench scale):
    narrow rows stage the addend rows in shared memory, wide rows prefetch them in registers; one, two and in-place
    addends, odd row ranges, must equal the plain result plus the addends bit for bit."""
    monkeypatch.setenv("IDDGCN_SPMM_V1", "0")
    from iddgcn_b200 import ops
    from iddgcn_b200.graph import Graph
    N, R = 70001, 2
    rng = np.random.default_rng(D)
    tri = np.stack([rng.integers(0, N, 400000), rng.integers(0, R, 400000), rng.integers(0, N, 400000)], 1).astype(np.int64)
    tri = tri[(tri[:, 2] % 11) != 4]                 # nodes without in-edges: rows that are pure addend
    g = Graph(tri, N, R)
    dS = torch.from_numpy(rng.standard_normal((R, N, D)).astype(np.float32)).cuda()
    a1 = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).cuda()
    a2 = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).cuda()
    plain = ops.spmm_bwd_x(g, dS)
    for lo, hi in [(0, N), (5, 40003), (40003, N)]:
        out = torch.full((N, D), -7.0, device="cuda")
        ops.spmm_bwd_x_rows(g, dS, out[lo:], lo, hi, addends=[a1[lo:]])
        assert torch.equal(out[lo:hi], plain[lo:hi] + a1[lo:hi])
        assert bool((out[:lo] == -7.0).all()) and bool((out[hi:] == -7.0).all())      # nothing outside the range
        ops.spmm_bwd_x_rows(g, dS, out[lo:], lo, hi, addends=[a1[lo:], a2[lo:]])
        assert torch.equal(out[lo:hi], plain[lo:hi] + (a1[lo:hi] + a2[lo:hi]))
        acc = a2.clone()
        ops.spmm_bwd_x_rows(g, dS, acc[lo:], lo, hi, addends=[acc[lo:]])
        assert torch.equal(acc[lo:hi], plain[lo:hi] + a2[lo:hi])
