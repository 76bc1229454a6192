mkdir -p gpurun_out/r2o
N=$1
for pos in first middle last; do
SOSWSOD_FC1_WGRAD_POS=$pos timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --blocks 3 > gpurun_out/r2o/bench${N}_$pos.json 2> gpurun_out/r2o/bench${N}_$pos.err
python - <<PY
import json
b=json.load(open("gpurun_out/r2o/bench${N}_$pos.json")); r=b["roofline"]
print("N=$N pos=$pos", round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["ms_per_step"],3), "gemm", round(r["gemm_ms_per_step"],3), "roi", round(r["roi_pool"]["fwd"]["ms_per_step"],3), round(r["roi_pool"]["bwd"]["ms_per_step"],3), "sgd", round(r["sgd_step"]["ms_per_step"],3))
PY
done
