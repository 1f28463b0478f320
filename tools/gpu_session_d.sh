#!/bin/bash
# round-2 GPU session D (1 GPU): whole GPU suite, smoke, host-step timeline, layer-level roofline, SDDMM U=1 variant
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/d; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" > $O/status.txt
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 600 python tools/debug_step_host.py --workload c3 > $O/timeline_c3_1gpu.json 2> $O/timeline.err; echo "timeline rc=$?" >> $O/status.txt
timeout 900 python tools/bench_layer.py > $O/layer.json 2> $O/layer.err; echo "layer rc=$?" >> $O/status.txt
timeout 300 python bench.py --workload c3 --no-e2e --no-cpu --check-rows 0 --steps 5 > $O/sddmm_default.json 2> $O/sddmm.err
IDDGCN_SDDMM_TUNE=2 timeout 300 python bench.py --workload c3 --no-e2e --no-cpu --check-rows 0 --steps 5 > $O/sddmm_u1.json 2>> $O/sddmm.err
tail -5 $O/pytest.log; cat $O/status.txt; tail -2 $O/smoke.log
