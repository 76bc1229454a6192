mkdir -p gpurun_out/r2n
N=$1
for mode in sharded allreduce; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/check_exchange.py $mode > gpurun_out/r2n/check${N}_$mode.json 2> gpurun_out/r2n/check${N}_$mode.err
  echo "== check N=$N $mode rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2n/check${N}_$mode.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('mode','n_gpus','max_rel_err_params','max_rel_err_momentum','bf16_operands_equal_cast_of_masters_on_every_rank','ok')})"
done
for mode in $2; do
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $mode > gpurun_out/r2n/bench${N}_$mode.json 2> gpurun_out/r2n/bench${N}_$mode.err
echo "== bench N=$N $mode rc=$?"; head -c 280 gpurun_out/r2n/bench${N}_$mode.json; echo; grep -v "NCCL INFO" gpurun_out/r2n/bench${N}_$mode.err | grep -i "error\|Traceback" -A8 | head -20
done
