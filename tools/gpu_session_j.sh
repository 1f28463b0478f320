#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/j; mkdir -p $O
timeout 600 python tools/bench_addends.py > $O/addends_minb3.json 2> $O/addends.err
IDDGCN_ADD_MINB=4 timeout 600 python tools/bench_addends.py > $O/addends_minb4.json 2>> $O/addends.err
cat $O/addends_minb3.json $O/addends_minb4.json
