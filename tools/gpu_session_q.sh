#!/bin/bash
# round-2 GPU session Q (1 GPU): ncu of the stream kernels at the per-rank shape of the 2x4 grid (32 floats per row)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/q; mkdir -p $O; : > $O/status.txt
CMD="python tools/bench_shapes.py --workload c4 --world 8 --grids 2x4 --iters 3"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_rows -s 4 -c 2 -o $O/prof_spmm_dc32 $CMD > $O/ncu.log 2>&1; echo "ncu rc=$?" >> $O/status.txt
cat $O/status.txt; tail -2 $O/plain.log | cut -c1-300
