#!/bin/bash
# round-2 GPU session B (1 GPU): re-measure after the addend-template fix; SDDMM variants; ncu on the c4 kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_graph_spmm.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" > $O/status.txt
timeout 600 python bench.py --workload c3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/status.txt
for k in 1 2 3 4; do IDDGCN_SDDMM_TUNE=$k timeout 300 python bench.py --workload c3 --no-e2e --no-cpu --check-rows 0 --steps 5 > $O/sddmm_tune$k.json 2>> $O/bench_c3.err; done
timeout 900 python bench.py > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?" >> $O/status.txt
timeout 900 python tools/bench_shapes.py --workload c4 --world 8 > $O/shapes_c4.json 2> $O/shapes_c4.err; echo "shapes c4 rc=$?" >> $O/status.txt
timeout 900 python tools/bench_shapes.py --workload c3 --world 8 > $O/shapes_c3x8.json 2> $O/shapes_c3x8.err; echo "shapes c3 rc=$?" >> $O/status.txt
CMD="python bench.py --no-e2e --no-cpu --check-rows 0 --steps 2 --warmup 3"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/status.txt
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_rows -s 6 -c 2 -o $O/prof_spmm_c4 $CMD > $O/ncu_spmm.log 2>&1; echo "ncu spmm rc=$?" >> $O/status.txt
$CMD > $O/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sddmm_stream -s 3 -c 1 -o $O/prof_sddmm_c4 $CMD > $O/ncu_sddmm.log 2>&1; echo "ncu sddmm rc=$?" >> $O/status.txt
tail -3 $O/pytest.log; cat $O/status.txt; ls -la $O
