#!/bin/bash
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z/tests.log
timeout 600 python bench.py --workload detect --images 1500 > gpurun_out/r2z/detect1.json 2> gpurun_out/r2z/detect1.err; echo "detect rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z/detect1.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['device_ms_per_image_rank_max'], d['generation_images_per_s'])
for k,v in d['stages'].items(): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='kernels' and a!='note'})
PY
tail -3 gpurun_out/r2z/detect1.err
