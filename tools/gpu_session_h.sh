#!/bin/bash
# round-2 GPU session H (1 GPU): whole GPU suite, smoke, layer-level roofline after the triple-stage changes, the driver's
# default bench command, ncu launch list + full capture of the final SDDMM and segment-scatter kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/h; mkdir -p $O; : > $O/status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 900 python tools/bench_layer.py > $O/layer.json 2> $O/layer.err; echo "layer rc=$?" >> $O/status.txt
IDDGCN_SEG_STREAM=0 timeout 900 python tools/bench_layer.py > $O/layer_old_scatter.json 2>> $O/layer.err
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/status.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" >> $O/status.txt
CMD="python bench.py --no-e2e --no-cpu --check-rows 0 --steps 2 --warmup 3"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'stream_rows|sddmm_stream|heavy_' -c 200 --csv --log-file $O/launches_c4.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/status.txt
$CMD > $O/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sddmm_stream -s 3 -c 1 -o $O/prof_sddmm_c4 $CMD > $O/ncu_sddmm.log 2>&1; echo "ncu sddmm rc=$?" >> $O/status.txt
CMD2="python tools/bench_layer.py --steps 1"
$CMD2 > $O/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'segment_stream|layer_combine' -s 6 -c 4 -o $O/prof_triple $CMD2 > $O/ncu_triple.log 2>&1; echo "ncu triple rc=$?" >> $O/status.txt
tail -4 $O/pytest.log; cat $O/status.txt; tail -1 $O/smoke.log
