#!/bin/bash
# final check of the tree (1 GPU): whole GPU suite, smoke, the driver's default bench command
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/s; mkdir -p $O; : > $O/status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/status.txt
tail -2 $O/pytest.log; cat $O/status.txt; tail -1 $O/smoke.log; cut -c1-400 $O/bench_default.json
