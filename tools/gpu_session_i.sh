#!/bin/bash
# round-2 GPU session I (2 GPUs): 2-GPU parity tests again (addend prefetch), phase timeline at the 2x4 per-rank shapes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/i; mkdir -p $O; : > $O/status.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
run() { name=$1; shift; timeout 900 python bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
run c4_dim32_2x1_stores_ABF --grid 2x1 --dim 32 --no-e2e --check-rows 200000
run c3x2_dim32_2x1_stores_ABF --workload c3 --grid 2x1 --dim 32 --no-e2e --check-rows 200000
run c4_2gpu_default
tail -3 $O/pytest.log; cat $O/status.txt
