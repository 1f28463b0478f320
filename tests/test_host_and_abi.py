"""CPU suite, part 2: host-side logic and the C-ABI surface (no compute calls — there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "iddgcn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(iddgcn_[a-z_0-9]+)\s*\(", hdr)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from iddgcn_b200 import _lib, build
    path = build.build()
    assert path.exists()
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/iddgcn.h but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature"
    out = subprocess.run(["nm", "-D", str(path)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T iddgcn_" in l}
    assert exported == set(declared), exported ^ set(declared)
    assert lib.iddgcn_version().decode().endswith("sm_100a")


def test_only_sm_100a_code_is_embedded():
    from iddgcn_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", str(build.LIB)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    from iddgcn_b200 import _lib
    from iddgcn_b200.graph import Graph
    from iddgcn_b200.model import get_IDDGCN_Model
    with pytest.raises(RuntimeError):
        Graph(np.zeros((1, 3), np.int64), 4, 1)
    with pytest.raises(RuntimeError):
        get_IDDGCN_Model(10, 2, 8, 8, seed=0)
    lib = _lib.load()
    h = C.c_void_p()
    tri = (C.c_int64 * 3)(0, 0, 1)
    rc = lib.iddgcn_graph_build(C.cast(tri, C.c_void_p), 1, 0, 4, 1, 0, 4, None, C.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.iddgcn_last_error()
    rc = lib.iddgcn_model_create(10, 2, 8, 3, C.byref(h))
    assert rc == -2


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: nothing inside the package does."""
    pkg = os.path.join(ROOT, "iddgcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            src = open(os.path.join(dirpath, fn)).read()
            m = re.search(r"^\s*(from|import)\s+oracle\b.*$", src, flags=re.M)
            assert m is None, f"{fn} imports oracle: {m.group(0)}"


def test_h5_reader_on_shipped_file_and_roundtrip(golden_dir, tmp_path):
    from iddgcn_b200 import h5
    path = os.path.join(golden_dir, "h5_sample_fold4.h5")
    layers = h5.read_keras_weights(path)
    with_w = [(n, ws) for n, ws in layers if ws]
    assert [n for n, _ in with_w] == ["entity_embeddings", "rgcn__layer", "rgcn__layer_1", "rgcn__layer_2", "DistMult"]
    assert [w for w, _ in with_w[1][1]] == ["relation_kernels:0", "self_kernel:0", "relation_weights:0", "W_alpha:0",
                                            "b_alpha:0"]
    ws = h5.read_weight_list(path)
    assert [w.shape for w in ws] == [(845, 64)] + [(4, 64, 64), (64, 64), (4,), (64, 4), (4,)] * 3 + [(4, 64)]
    assert sum(w.size for w in ws) == 116568
    z = np.load(os.path.join(golden_dir, "weights_fold4.npz"))
    for k, w in enumerate(ws):
        assert np.array_equal(w, z[f"w{k:02d}"])
    out = tmp_path / "rt.h5"
    h5.write_keras_weights(out, layers)
    back = h5.read_keras_weights(out)
    assert [n for n, _ in back] == [n for n, _ in layers]
    for (_, a), (_, b) in zip(layers, back):
        assert [n for n, _ in a] == [n for n, _ in b]
        for (_, x), (_, y) in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y)


def test_bench_algorithmic_bytes_match_survey():
    import bench
    # C3 B_f = 11.35 GB, B_b = 10.84 GB -> 1109.6 B/edge; C4 123.84 + 113.60 GB -> 1187.2 B/edge
    bf, bb = bench.algorithmic_bytes(20_000_000, 1_000_000, 1_000_000, 2, 128)
    assert round(bf / 1e9, 2) == 11.35 and round(bb / 1e9, 2) == 10.84
    assert round((bf + bb) / 20_000_000, 1) == 1109.6
    bf, bb = bench.algorithmic_bytes(200_000_000, 20_000_000, 20_000_000, 2, 128)
    assert round(bf / 1e9, 2) == 123.84 and round(bb / 1e9, 2) == 113.60
    assert round((bf + bb) / 200_000_000, 1) == 1187.2


def test_bench_cpu_inputs_are_a_valid_graph():
    """The CPU arm's structure is the CSR/CSC of the SAME raw edge list the GPU arm builds its graph from
    (bench.raw_edges), and the two feature generators (NumPy for the CPU arm / the parity check, torch for the device)
    agree bit for bit."""
    import torch
    import bench
    from oracle import iddgcn_oracle as O
    wl = dict(bench.WORKLOADS["small"]); wl["name"] = "small"; wl["N"] = 2000; wl["per_rel"] = 15000
    struct, N, nnz, X, dS, sample = bench.cpu_sample_inputs(wl)
    tri = []
    for r in range(wl["R"]):
        h, t = bench.raw_edges(N, wl["per_rel"], r)
        tri.append(np.stack([h, np.full_like(h, r), t], 1))
    ref = O.csr_csc_from_adj(O.get_adj_mats(np.concatenate(tri), N, wl["R"]), N)
    for k in struct:
        assert np.array_equal(struct[k], ref[k]), k
    assert nnz == ref["nnz"]
    ids = np.array([0, 1, 77, N - 1, 2 * N - 1], np.int64)
    for kind, full in (("X", X), ("dS", dS)):
        a = bench.gen_rows_np(kind, ids, 32, 64, wl["D"])
        b = bench.gen_rows_torch(torch, kind, torch.from_numpy(ids), 32, 64, wl["D"]).numpy()
        assert np.array_equal(a, b)
        ok = ids < full.shape[0]
        assert np.array_equal(full[ids[ok], 32:96], a[ok])
    assert 0.45 < X.mean() < 0.55 and abs(dS.mean()) < 0.05 and 0.9 < dS.std() < 1.1


def test_product_fold_pipeline_matches_oracle_and_shipped_files(golden_dir):
    """tools/train_folds.fold_frames (host data prep of config 2) == oracle.fold_pipeline, which is pinned to the
    shipped fold files by sha256 (test_oracle_pins)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_folds
    from oracle import iddgcn_oracle as O
    mine = train_folds.fold_frames(os.path.join(golden_dir, "data"))
    ref = O.fold_pipeline(os.path.join(golden_dir, "data"))
    for a, b in zip(mine, ref):
        assert np.array_equal(a["X_train"], b["X_train"].values)
        assert np.array_equal(a["X_test"], b["X_test"].values)
        assert np.array_equal(a["neg_train"], b["X_train_neg"][0].astype(np.int64))
        assert np.array_equal(a["neg_test"], b["neg_X_test"].values.astype(np.int64))


def test_product_ground_truth_and_metrics_match_oracle(golden_dir, tmp_path):
    """iddgcn_b200.explain_eval (table look-ups) against the loop restatement in the oracle and the shipped gt_all.csv,
    then Precision/Recall/F1@5 on synthetic predictions that mix direct hits, reversed hits and misses."""
    import pandas as pd
    from iddgcn_b200 import explain_eval as E
    from oracle import iddgcn_oracle as O
    d = os.path.join(golden_dir, "data")
    dc, cc, dd = (pd.read_csv(os.path.join(d, fn)).values
                  for fn in ("triplets_dc.csv", "mu_similar0.97.csv", "drug_similar0.78.csv"))
    rows = E.construct_ground_truth(dc, cc, dd)
    z = np.load(os.path.join(golden_dir, "gt_all.npz"))
    assert rows == [z["flat"][z["offsets"][i]:z["offsets"][i + 1]].tolist() for i in range(len(z["offsets"]) - 1)]
    sub = slice(0, 300)                                                         # the loop oracle on a slice (seconds)
    assert E.construct_ground_truth(dc[sub], cc, dd) == O.construct_ground_truth(dc[sub], cc, dd)

    test_triples = pd.read_csv(os.path.join(d, "test_filtered_fold4.csv")).values
    extra = np.array([[0, 0, 1], [844, 1, 2]])                                  # triples without ground truth drop out
    kept, gts = E.filter_ground_truth(np.concatenate([extra[:1], test_triples, extra[1:]]), rows)
    kept_o, gts_o = O.filter_ground_truth(np.concatenate([extra[:1], test_triples, extra[1:]]), rows)
    assert np.array_equal(kept, test_triples) and np.array_equal(kept_o, kept) and gts == gts_o

    rng = np.random.default_rng(0)
    preds = []
    for g in gts:
        g = np.asarray(g, np.int64)
        n = rng.integers(1, 11)
        fake = rng.integers(0, 845, (n, 3))
        fake[:, 1] = rng.integers(0, 4, n)
        take = g[rng.integers(0, len(g), rng.integers(0, 6))]
        if rng.random() < 0.5:
            take = take[:, ::-1]
        p = np.concatenate([take, fake])[:10]
        preds.append(p[rng.permutation(len(p))])
    mine, ref = E.explanation_metrics(preds, gts), O.explanation_metrics(preds, gts_o)
    assert mine == ref and len(mine) == 288
    avg = E.average_metrics(mine)
    assert avg["Precision@5"] == pytest.approx(np.mean([m[0] for m in ref]))
    E.save_ground_truth(tmp_path / "gt.npz", kept, gts)
    kept2, gts2 = E.load_ground_truth(tmp_path / "gt.npz")
    assert np.array_equal(kept2, kept) and gts2 == gts


def test_ground_truth_loader_reads_the_reference_container(tmp_path):
    """gt_process.py:39 stores the ground truth as a pickled object array under 'result'; the loader opens it only when
    the caller opts in, and refuses (NumPy's ValueError) otherwise."""
    from iddgcn_b200 import explain_eval as E
    gt = [[[1.0, 0.0, 700.0], [700.0, 2.0, 701.0]], [[5.0, 1.0, 690.0]]]
    path = tmp_path / "gt_filtered_fold9.npz"
    np.savez(path, result=np.array(gt, dtype=object))
    with pytest.raises(ValueError):
        E.load_ground_truth(path)
    triples, back = E.load_ground_truth(path, allow_pickle=True)
    assert triples is None and back == gt
    preds = [np.array([[1, 0, 700], [9, 9, 9]]), np.array([[690, 1, 5]])]
    (p0, r0, _), (p1, r1, _) = E.explanation_metrics(preds, back)
    assert (p0, r0) == (0.2, 0.1) and (p1, r1) == (0.2, 0.1)
