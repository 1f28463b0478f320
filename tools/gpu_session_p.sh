#!/bin/bash
# round-2 GPU session P (4 GPUs): e2e copy-schedule selection under PCIe contention, p2p exchange with two column groups
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/p; mkdir -p $O; : > $O/status.txt
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/pytest_dist.log 2>&1; echo "pytest dist rc=$?" >> $O/status.txt
run() { name=$1; n=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
run c4_4gpu_default 4
run c3x4_4gpu_2x2 4 --workload c3 --grid 2x2 --check-rows 400000
tail -2 $O/pytest_dist.log; cat $O/status.txt
