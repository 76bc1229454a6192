mkdir -p gpurun_out/r2t
N=$1
for steps in 1 3; do
  CHECK_STEPS=$steps timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/check_exchange.py nvls > gpurun_out/r2t/check${N}_nvls_$steps.json 2> gpurun_out/r2t/check${N}_nvls_$steps.err
  echo "== check N=$N nvls steps=$steps rc=$?"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2t/check${N}_nvls_$steps.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('mode','n_gpus','max_rel_err_params','max_rel_err_momentum','bf16_operands_equal_cast_of_masters_on_every_rank','ok')})
except Exception as e: print('no json', e)"
  grep -v "NCCL INFO" gpurun_out/r2t/check${N}_nvls_$steps.err | grep -i "error\|Traceback" -B2 -A12 | head -30
done
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2t/bench${N}_auto.json 2> gpurun_out/r2t/bench${N}_auto.err
python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2t/bench${N}_auto.json")); r=b["roofline"]
    print("N=$N auto ->", b["exchange_check"]["mode"], round(b["value"],1), round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["value"],1), round(b["e2e"]["ms_per_step"],3), "gemm", round(r["gemm_ms_per_step"],3), "roi", round(r["roi_pool"]["fwd"]["ms_per_step"],3), round(r["roi_pool"]["bwd"]["ms_per_step"],3), "nvls", r.get("nvls_update",{}).get("ms_per_step"), b["exchange_check"]["ok"])
except Exception as e: print("no json", e)
PY
grep -v "NCCL INFO" gpurun_out/r2t/bench${N}_auto.err | grep -i "error\|Traceback\|unavailable" -A10 | head -20
