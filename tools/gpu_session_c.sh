#!/bin/bash
# round-2 GPU session C (2 GPUs): fused peer-memory exchange — parity tests, then bench lines for the 2-GPU grids
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/c; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; free -g >> $O/topo.txt; nproc >> $O/topo.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" > $O/status.txt
run() { name=$1; shift; timeout 900 python bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
run c4_2x1_p2p --grid 2x1
run c4_1x2 --grid 1x2 --no-e2e
run c4_2x1_nccl --grid 2x1 --exchange nccl --no-e2e
run c3x2_2x1_p2p --workload c3 --grid 2x1
run c3x2_1x2 --workload c3 --grid 1x2 --no-e2e
tail -5 $O/pytest.log; cat $O/status.txt
