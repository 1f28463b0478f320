#!/bin/bash
# round-2 GPU session A (1 GPU): new kernels' parity, bench lines for small/c3/c4, per-rank shape sweep, SDDMM variants
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/a; mkdir -p $O
{ free -g; nproc; cat /sys/fs/cgroup/memory.max 2>/dev/null; nvidia-smi --query-gpu=name,memory.total --format=csv; } > $O/box.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_graph_spmm.py tests/test_bench_contract.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/box.txt
timeout 300 python bench.py --workload small --steps 5 --warmup 3 > $O/bench_small.json 2> $O/bench_small.err; echo "small rc=$?" >> $O/box.txt
timeout 600 python bench.py --workload c3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/box.txt
for k in 1 2 3 4; do IDDGCN_SDDMM_TUNE=$k timeout 300 python bench.py --workload c3 --no-e2e --no-cpu --check-rows 0 --steps 5 > $O/sddmm_tune$k.json 2>> $O/bench_c3.err; done
timeout 900 python bench.py > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?" >> $O/box.txt
timeout 900 python tools/bench_shapes.py --workload c4 --world 8 > $O/shapes_c4.json 2> $O/shapes_c4.err; echo "shapes c4 rc=$?" >> $O/box.txt
timeout 900 python tools/bench_shapes.py --workload c3 --world 8 > $O/shapes_c3x8.json 2> $O/shapes_c3x8.err; echo "shapes c3 rc=$?" >> $O/box.txt
tail -3 $O/pytest.log; cat $O/box.txt
