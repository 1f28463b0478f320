#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/o; mkdir -p $O; : > $O/status.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python tools/bench_layer.py > $O/layer.json 2> $O/layer.err; echo "layer rc=$?" >> $O/status.txt
tail -3 $O/pytest.log; cat $O/status.txt
python -c "
import json
d=json.loads(open('$O/layer.json').read().strip().splitlines()[-1]); g=d['roofline_by_group']['groups']
print('ms/step',d['ms_per_step'],'dW_alpha',g['dW_alpha']['ms'],g['dW_alpha']['hbm_frac'])"
