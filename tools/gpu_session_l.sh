#!/bin/bash
# round-2 GPU session L (1 GPU): final state — whole GPU suite, smoke, default bench, skewed (hub-row) variant
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/l; mkdir -p $O; : > $O/status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 600 python bench.py --workload c3zipf > $O/bench_c3zipf.json 2> $O/bench_c3zipf.err; echo "zipf rc=$?" >> $O/status.txt
timeout 600 python bench.py --workload c3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/status.txt
tail -3 $O/pytest.log; cat $O/status.txt; tail -1 $O/smoke.log
