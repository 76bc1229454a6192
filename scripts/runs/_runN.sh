mkdir -p gpurun_out/r2v
N=$1
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v/bench${N}_auto.json 2> gpurun_out/r2v/bench${N}_auto.err
python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2v/bench${N}_auto.json")); r=b["roofline"]
    print("N=$N auto ->", b["exchange_check"]["mode"], round(b["value"],1), round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["value"],1), round(b["e2e"]["ms_per_step"],3), "nvls", r.get("nvls_update",{}).get("ms_per_step"), b["exchange_check"]["ok"])
except Exception as e: print("no json", e)
PY
grep -v "NCCL INFO" gpurun_out/r2v/bench${N}_auto.err | grep -i "error\|Traceback\|unavailable" -A10 | head -20
