#!/usr/bin/env python3
"""Bandwidth of the hand-written multicast exchange kernels (csrc/nvls.cu) on the ranks of one box:
all-gather leg (multimem.st of one 512 MB block per rank), reduce-scatter leg (multimem.ld_reduce.add of one block
per rank), and both concurrently on two streams, for several CTA counts.
    torchrun --nproc-per-node P tools/nvls_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iddgcn_b200._lib import check, load  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
lib = load()
n = 1_000_000 * 128                       # floats per block (512 MB)
full = symm_mem.empty(n * world, dtype=torch.float32, device=dev)
part = symm_mem.empty(n * world, dtype=torch.float32, device=dev)
hf = symm_mem.rendezvous(full, dist.group.WORLD)
hp = symm_mem.rendezvous(part, dist.group.WORLD)
mcf, mcp = int(hf.multicast_ptr), int(hp.multicast_ptr)
assert mcf and mcp, "no multicast support"
loc = torch.rand(n, device=dev)
part.copy_(torch.rand(n * world, device=dev))
out = torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=-1)
off = rank * n * 4


def ag(ctas, st):
    check(lib.iddgcn_mc_store(mcf + off, loc.data_ptr(), n, ctas, st.cuda_stream))


def rs(ctas, st):
    check(lib.iddgcn_mc_reduce_load(mcp + off, out.data_ptr(), n, ctas, st.cuda_stream))


def timed(fn, iters=5):
    cur = torch.cuda.current_stream()
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); hf.barrier(channel=1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    cur.wait_stream(s1); cur.wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def on(st, f, ctas):
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        f(ctas, st)


res = {"world": world, "block_bytes": n * 4}
for ctas in (32, 64, 128, 256):
    t_ag = timed(lambda: on(s1, ag, ctas))
    t_rs = timed(lambda: on(s2, rs, ctas))
    t_both = timed(lambda: (on(s1, ag, ctas), on(s2, rs, ctas)))
    res[f"ctas{ctas}"] = {"all_gather_ms": t_ag, "reduce_scatter_ms": t_rs, "both_ms": t_both,
                          "ag_recv_GBps": (world - 1) * n * 4 / t_ag / 1e6, "rs_send_GBps": (world - 1) * n * 4 / t_rs / 1e6}
# correctness of both legs
hf.barrier(channel=1); torch.cuda.synchronize()
ag(64, torch.cuda.current_stream()); torch.cuda.synchronize(); hf.barrier(channel=1); torch.cuda.synchronize()
gathered = [torch.empty_like(loc) for _ in range(world)]
dist.all_gather(gathered, loc)
res["all_gather_exact"] = bool(torch.equal(full, torch.cat(gathered)))
rs(64, torch.cuda.current_stream()); torch.cuda.synchronize()
ref = part.clone(); dist.all_reduce(ref)
res["reduce_scatter_max_rel_err"] = float(((out - ref[rank * n:(rank + 1) * n]).abs().max() / ref.abs().max()).item())
if rank == 0:
    print(json.dumps(res), flush=True)
hf.barrier(channel=1); torch.cuda.synchronize()
dist.destroy_process_group()
