#!/usr/bin/env python3
"""Measures, on the ranks of one box, the exchange primitives of the partitioned SpMM: all-gather alone,
reduce-scatter alone, and both concurrently on two communicators (block = 1M rows x 128 fp32 = 512 MB per rank).
    torchrun --nproc-per-node P tools/nccl_probe.py
"""
import json
import os

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
opts = dist.ProcessGroupNCCL.Options(); opts.is_high_priority_stream = True
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
g2 = dist.new_group(backend="nccl", pg_options=opts)
n = 1_000_000 * 128
loc = torch.rand(n, device=dev); full = torch.empty(n * world, device=dev)
part = torch.rand(n * world, device=dev); out = torch.empty(n, device=dev)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def both():
    w1 = dist.all_gather_into_tensor(full, loc, async_op=True)
    w2 = dist.reduce_scatter_tensor(out, part, group=g2, async_op=True)
    w1.wait(); w2.wait()


res = {"world": world, "NCCL_ALGO": os.environ.get("NCCL_ALGO"), "NCCL_PROTO": os.environ.get("NCCL_PROTO"),
       "bytes_per_rank_block": n * 4,
       "all_gather_ms": timed(lambda: dist.all_gather_into_tensor(full, loc)),
       "reduce_scatter_ms": timed(lambda: dist.reduce_scatter_tensor(out, part, group=g2)),
       "both_concurrent_ms": timed(both)}
res["recv_GBps_all_gather"] = (world - 1) * n * 4 / res["all_gather_ms"] / 1e6
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
