mkdir -p gpurun_out
for k in 0 1 2 3 4 5 6 7 8 9 10 11 12; do
  IDDGCN_TUNE=$k timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/tune_$k.log 2>&1
  tail -1 gpurun_out/tune_$k.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('tune $k', round(d['value']/1e9,3), round(r['fwd']['ms'],4), round(r['bwd']['ms'],4), d['clocks']['sm_mhz'])" || tail -2 gpurun_out/tune_$k.log
done
