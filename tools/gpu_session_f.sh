#!/bin/bash
# round-2 GPU session F (8 GPUs): the driver's scaling configuration and its alternatives; configs[1] and [4] at full width
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/f; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; free -g >> $O/topo.txt; nproc >> $O/topo.txt
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 "$@"; }
run() { name=$1; n=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/status.txt; }
: > $O/status.txt
run c4_8gpu_default 8
run c3x8_8gpu_default 8 --workload c3
run c4_4gpu_default 4
run c4_8gpu_2x4_dma_AFB 8 --partials dma --order AFB --no-e2e --check-rows 200000
run c4_8gpu_1x8 8 --grid 1x8 --no-e2e --check-rows 200000
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 tools/debug_step_host.py --workload c4 --grid 2x4 > $O/timeline_c4_2x4.jsonl 2> $O/timeline.err; echo "timeline rc=$?" >> $O/status.txt
# configs[1]: five-fold CV, folds data-parallel on 5 GPUs; configs[4]: 10k explanation triples on 8 GPUs
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv,noheader,nounits -lms 250 > $O/busy_folds.csv &
SMI=$!
timeout 900 python tools/train_folds.py --gpus 5 --epochs 5000 > $O/five_fold_cv_5gpu_5000epochs.jsonl 2> $O/folds.err; echo "folds rc=$?" >> $O/status.txt
kill $SMI
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv,noheader,nounits -lms 250 > $O/busy_explain.csv &
SMI=$!
timeout 900 python tools/bench_explain.py --gpus 8 --triples 10000 > $O/explain_c5_10k_triples_8gpu.json 2> $O/explain.err; echo "explain rc=$?" >> $O/status.txt
kill $SMI
run c4_8gpu_4x2 8 --grid 4x2 --no-e2e --check-rows 200000
cat $O/status.txt
