#!/bin/bash
mkdir -p gpurun_out/r2x
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2x/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2x/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x/tests.log
timeout 600 python bench.py > gpurun_out/r2x/bench1.json 2> gpurun_out/r2x/bench1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x/bench1.json').read().strip().splitlines()[-1]); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],3), [round(x,2) for x in d['blocks_ms_per_step']], 'e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],3), 'sgd', r['sgd_step']['ms_per_step'], r['sgd_step'].get('overlapped_in_timed_blocks'), 'parity', d.get('parity_checked'), d.get('parity'), 'cpu', d['cpu_baseline']['value'], d['clocks']['sm_mhz'], d['gpu_launches'])
PY
timeout 600 python bench.py --workload detect --images 1500 > gpurun_out/r2x/detect1.json 2> gpurun_out/r2x/detect1.err; echo "detect rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x/detect1.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['device_ms_per_image_rank_max'])
for k,v in d['stages'].items(): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
tail -5 gpurun_out/r2x/detect1.err
