#!/bin/bash
mkdir -p gpurun_out/r2z
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2z/bench2.json 2> gpurun_out/r2z/bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2z/bench2.json").read().strip().splitlines()[-1]); r=b["roofline"]
print("N=2", b["exchange_check"]["mode"], round(b["value"],1), round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["value"],1), "ok", b["exchange_check"]["ok"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload detect --images 2000 > gpurun_out/r2z/detect2.json 2> gpurun_out/r2z/detect2.err; echo "detect2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z/detect2.json').read().strip().splitlines()[-1])
print("detect N=2", round(d['value'],1), d['device_ms_per_image_rank_max'], d['generation_images_per_s'], d['gather_and_json_dump_s'])
PY
grep -v "NCCL INFO" gpurun_out/r2z/bench2.err | grep -i "error\|Traceback" -A8 | head -20
