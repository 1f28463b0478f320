#!/usr/bin/env python3
"""Dense transform timing: fp32 SIMT kernels against the tcgen05 3xTF32 kernel at the layer-level line's shape
(M = 2*(2^22 + 2^19) gathered rows, D = 128) and at D = 64.

    python tools/bench_gemm.py [--rows 9437184] [--iters 10]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2 * (2 ** 22 + 2 ** 19))
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    import torch
    from iddgcn_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    out = {"rows": args.rows}
    for D in (128, 64):
        M = args.rows
        a = torch.randn(M, D, device=dev)
        w = torch.randn(D, D, device=dev) / D ** 0.5
        c = torch.empty(M, D, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        for name, path in (("simt_fp32", 1), ("tcgen05_tf32x3", 2)):
            for acc in (0, 1):
                call = lambda: _lib.check(lib.iddgcn_dense_transform(_lib.ptr(a), _lib.ptr(w), M, D, 0, _lib.ptr(c), acc,
                                                                     path, C.c_void_p(st)))
                for _ in range(3):
                    call()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(args.iters):
                    call()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                out[f"D{D}_{name}_{'accumulate' if acc else 'overwrite'}"] = {
                    "ms": round(ms, 4), "TFLOPs": round(2.0 * M * D * D / ms / 1e9, 2),
                    "GBps": round((2 + acc) * M * D * 4 / ms / 1e6, 1)}
        dw = torch.empty(D, D, device=dev)
        for name, path in (("simt_fp32", 1), ("tcgen05_tf32x3", 2)):
            call = lambda: _lib.check(lib.iddgcn_dense_transform_grad(_lib.ptr(a), _lib.ptr(c), M, D, _lib.ptr(dw), 0, path,
                                                                      C.c_void_p(st)))
            for _ in range(3):
                call()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.iters):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            out[f"D{D}_grad_{name}"] = {"ms": round(ms, 4), "TFLOPs": round(2.0 * M * D * D / ms / 1e9, 2),
                                        "GBps": round(2 * M * D * 4 / ms / 1e6, 1)}
        ref = (a[:4096].double() @ w.double()).float()
        _lib.check(lib.iddgcn_dense_transform(_lib.ptr(a), _lib.ptr(w), 4096, D, 0, _lib.ptr(c), 0, 2, C.c_void_p(st)))
        torch.cuda.synchronize()
        out[f"D{D}_tc_max_abs_err_vs_fp64"] = float((c[:4096] - ref).abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
