#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/n; mkdir -p $O; : > $O/status.txt
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_bench_contract.py tests/test_gpu_model.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python bench.py --workload c3zipf > $O/bench_c3zipf.json 2> $O/bench_c3zipf.err; echo "zipf rc=$?" >> $O/status.txt
timeout 900 python tools/bench_layer.py > $O/layer.json 2> $O/layer.err; echo "layer rc=$?" >> $O/status.txt
tail -3 $O/pytest.log; cat $O/status.txt
