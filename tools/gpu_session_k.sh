#!/bin/bash
# round-2 GPU session K (8 GPUs): final confirmation of the driver's scaling command after the addend-epilogue fix
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/k; mkdir -p $O; : > $O/status.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_graph_spmm.py -x -q -m gpu -k "addends or row_ranges" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
run() { name=$1; n=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
run c4_8gpu_final 8
run c3x8_8gpu_final 8 --workload c3
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/pytest_dist.log 2>&1; echo "pytest dist rc=$?" >> $O/status.txt
tail -2 $O/pytest.log; tail -2 $O/pytest_dist.log; cat $O/status.txt
