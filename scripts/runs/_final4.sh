#!/bin/bash
mkdir -p gpurun_out/r2z
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --workload detect --images 5000 > gpurun_out/r2z/detect4.json 2> gpurun_out/r2z/detect4.err; echo "detect4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z/detect4.json').read().strip().splitlines()[-1])
print("detect N=4", round(d['value'],1), d['device_ms_per_image_rank_max'], d['generation_images_per_s'], d['gather_and_json_dump_s'], d['detection_rows_written'])
PY
grep -i "error\|Traceback" -A8 gpurun_out/r2z/detect4.err | head -20
