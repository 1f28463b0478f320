#!/bin/bash
# round-2 GPU session R (1 GPU): launch-shape sweep of the stream kernel at 32 floats per row (per-rank shape of 2x4)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r; mkdir -p $O
for k in 0 3 6 8 13 14 2; do
  IDDGCN_TUNE=$k timeout 300 python tools/bench_shapes.py --workload c4 --world 8 --grids 2x4 --iters 5 2> $O/err_$k.txt | head -1 > $O/tune_$k.json
  echo "tune $k: $(python -c "import json;d=json.load(open('$O/tune_$k.json'));print(d['fwd_ms'],d['bwd_ms'],d['fwd_frac'])")"
done
