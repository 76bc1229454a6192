mkdir -p gpurun_out/r2q
N=$1
for v in async main; do
if [ $v = main ]; then export SOSWSOD_UPDATE_ON_MAIN=1; else unset SOSWSOD_UPDATE_ON_MAIN; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --blocks 3 --exchange nvls > gpurun_out/r2q/bench${N}_$v.json 2> gpurun_out/r2q/bench${N}_$v.err
python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2q/bench${N}_$v.json")); r=b["roofline"]
    print("N=$N nvls $v", round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["ms_per_step"],3), "gemm", round(r["gemm_ms_per_step"],3), "roi", round(r["roi_pool"]["fwd"]["ms_per_step"],3), round(r["roi_pool"]["bwd"]["ms_per_step"],3), "sgd", round(r["sgd_step"]["ms_per_step"],3), "nvls", r.get("nvls_update"))
except Exception as e: print("no json", e)
PY
grep -v "NCCL INFO" gpurun_out/r2q/bench${N}_$v.err | grep -i "error\|Traceback" -A10 | head -20
done
