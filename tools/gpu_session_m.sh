#!/bin/bash
# round-2 GPU session M (1 GPU): A/B of the tensor-core transform with A_hi supplied by the hardware's own truncation
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/m; mkdir -p $O; : > $O/status.txt
timeout 600 python -m pytest tests/test_gpu_dense.py -x -q -m gpu > $O/pytest_default.log 2>&1; echo "dense default rc=$?" >> $O/status.txt
IDDGCN_TC_AHI_TRUNC=1 timeout 600 python -m pytest tests/test_gpu_dense.py -x -q -m gpu > $O/pytest_trunc.log 2>&1; echo "dense trunc rc=$?" >> $O/status.txt
timeout 600 python tools/bench_gemm.py > $O/gemm_default.json 2> $O/gemm.err; echo "gemm default rc=$?" >> $O/status.txt
IDDGCN_TC_AHI_TRUNC=1 timeout 600 python tools/bench_gemm.py > $O/gemm_trunc.json 2>> $O/gemm.err; echo "gemm trunc rc=$?" >> $O/status.txt
tail -3 $O/pytest_trunc.log; cat $O/status.txt; cat $O/gemm_default.json $O/gemm_trunc.json
