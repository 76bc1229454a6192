mkdir -p gpurun_out/r2p
N=$1
for mode in nvls sharded; do
  for steps in 1 3; do
  CHECK_STEPS=$steps timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/check_exchange.py $mode > gpurun_out/r2p/check${N}_${mode}_$steps.json 2> gpurun_out/r2p/check${N}_${mode}_$steps.err
  echo "== check N=$N $mode steps=$steps rc=$?"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2p/check${N}_${mode}_$steps.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('mode','n_gpus','max_rel_err_params','max_rel_err_momentum','bf16_operands_equal_cast_of_masters_on_every_rank','ok')})
except Exception as e: print('no json', e)"
  grep -v "NCCL INFO" gpurun_out/r2p/check${N}_${mode}_$steps.err | grep -i "error\|Traceback" -B2 -A12 | head -40
  done
done
for mode in $2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --blocks 3 --exchange $mode > gpurun_out/r2p/bench${N}_$mode.json 2> gpurun_out/r2p/bench${N}_$mode.err
echo "== bench N=$N $mode rc=$?"; python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2p/bench${N}_$mode.json")); r=b["roofline"]
    print("N=$N $mode", round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "e2e", round(b["e2e"]["ms_per_step"],3), "gemm", round(r["gemm_ms_per_step"],3), "roi", round(r["roi_pool"]["fwd"]["ms_per_step"],3), round(r["roi_pool"]["bwd"]["ms_per_step"],3), "sgd", round(r["sgd_step"]["ms_per_step"],3), b["exchange_check"])
except Exception as e: print("no json", e)
PY
grep -v "NCCL INFO" gpurun_out/r2p/bench${N}_$mode.err | grep -i "error\|Traceback\|nvls exchange unavailable" -A10 | head -30
done
