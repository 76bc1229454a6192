mkdir -p gpurun_out/r2s
N=$1
for c in $2; do
SOSWSOD_NVLS_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --blocks 3 --exchange nvls > gpurun_out/r2s/bench${N}_c$c.json 2> gpurun_out/r2s/bench${N}_c$c.err
python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2s/bench${N}_c$c.json")); r=b["roofline"]
    print("N=$N nvls ctas=$c", round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["ms_per_step"],3), [round(x,2) for x in b["e2e"]["blocks_ms_per_step"]], "gemm", round(r["gemm_ms_per_step"],3), "roi", round(r["roi_pool"]["fwd"]["ms_per_step"],3), round(r["roi_pool"]["bwd"]["ms_per_step"],3), "nvls", round(r["nvls_update"]["ms_per_step"],3))
except Exception as e: print("no json", e)
PY
grep -v "NCCL INFO" gpurun_out/r2s/bench${N}_c$c.err | grep -i "error\|Traceback" -A10 | head -20
done
