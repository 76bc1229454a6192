#!/bin/bash
mkdir -p gpurun_out/r2z
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload detect --images 5000 > gpurun_out/r2z/detect8.json 2> gpurun_out/r2z/detect8.err; echo "detect8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z/detect8.json').read().strip().splitlines()[-1])
print("detect N=8", round(d['value'],1), d['device_ms_per_image_rank_max'], d['generation_images_per_s'], d['gather_and_json_dump_s'], d['detection_rows_written'])
PY
grep -i "error\|Traceback" -A8 gpurun_out/r2z/detect8.err | head -20
