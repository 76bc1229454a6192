#!/bin/bash
# A/B of the single-GPU update placement: under the backward (default) vs after the step
mkdir -p gpurun_out/r2w
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "sgd or plugin_surface or hook" > gpurun_out/r2w/tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2w/tests.log
for mode in over after over after; do
  if [ $mode = after ]; then export SOSWSOD_NO_OVERLAP_UPDATE=1; else unset SOSWSOD_NO_OVERLAP_UPDATE; fi
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2w/bench_${mode}_$SECONDS.json 2> gpurun_out/r2w/bench_${mode}.err
  python - <<PY
import json,glob
f=sorted(glob.glob('gpurun_out/r2w/bench_${mode}_*.json'))[-1]
d=json.loads(open(f).read().strip().splitlines()[-1])
r=d['roofline']
print('$mode', round(d['value'],1), round(d['ms_per_step'],3), [round(x,2) for x in d['blocks_ms_per_step']], 'e2e', round(d['e2e']['value'],1), 'gemm frac', round(r['frac'],3), 'sgd', round(r['sgd_step']['ms_per_step'],3), r['sgd_step'].get('overlapped'), 'roi', {k:round(v['ms_per_step'],3) for k,v in r['roi_pool'].items() if isinstance(v,dict) and 'ms_per_step' in v}, d['clocks']['sm_mhz'])
PY
done
