#!/bin/bash
# round-2 GPU session G (2 GPUs): phase timelines of the p2p step at the per-rank shapes of the 2x4 grid (D=32, 2x1),
# host-step timeline after the push moved to its own stream
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/g; mkdir -p $O; : > $O/status.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { name=$1; shift; timeout 900 python bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
run c4_dim32_2x1_stores_ABF --grid 2x1 --dim 32 --no-e2e --check-rows 200000
run c4_dim32_2x1_dma_AFB --grid 2x1 --dim 32 --partials dma --order AFB --no-e2e --check-rows 200000
run c4_dim32_2x1_stores_AFB --grid 2x1 --dim 32 --order AFB --no-e2e --check-rows 200000
timeout 300 $TR tools/debug_step_host.py --workload c3 --grid 2x1 > $O/timeline_c3x2_2x1.json 2> $O/timeline_2x1.err; echo "tl 2x1 rc=$?" >> $O/status.txt
run c3x2_2x1_e2e --workload c3 --grid 2x1 --check-rows 200000
cat $O/status.txt
