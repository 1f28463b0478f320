#!/bin/bash
# round-2 GPU session E (2 GPUs): host-step timelines at N=2, exchange schedule variants
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/e; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/debug_step_host.py --workload c3 --grid 2x1 > $O/timeline_c3x2_2x1.json 2> $O/timeline_2x1.err; echo "tl 2x1 rc=$?" > $O/status.txt
timeout 300 $TR tools/debug_step_host.py --workload c3 --grid 1x2 > $O/timeline_c3x2_1x2.json 2> $O/timeline_1x2.err; echo "tl 1x2 rc=$?" >> $O/status.txt
run() { name=$1; shift; timeout 900 python bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
run c4_2x1_dma_AFB --grid 2x1 --partials dma --no-e2e --check-rows 200000
run c4_2x1_stores_ABF --grid 2x1 --order ABF --no-e2e --check-rows 200000
run c4_2x1_dma_ABF --grid 2x1 --partials dma --order ABF --no-e2e --check-rows 200000
run c3x2_2x1_dma_AFB --workload c3 --grid 2x1 --partials dma --no-e2e --check-rows 200000
run c3x2_1x2_e2e --workload c3 --grid 1x2 --check-rows 200000
cat $O/status.txt
