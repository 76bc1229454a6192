mkdir -p gpurun_out/r2m
N=$1
for mode in sharded allreduce; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/check_exchange.py $mode > gpurun_out/r2m/check${N}_$mode.json 2> gpurun_out/r2m/check${N}_$mode.err
  echo "== check N=$N $mode rc=$?"; cat gpurun_out/r2m/check${N}_$mode.json; grep -v "NCCL INFO" gpurun_out/r2m/check${N}_$mode.err | grep -i "error\|Traceback" -A8 | head -30
done
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m/bench${N}_sharded.json 2> gpurun_out/r2m/bench${N}_sharded.err
echo "== bench N=$N rc=$?"; head -c 300 gpurun_out/r2m/bench${N}_sharded.json; echo; grep -v "NCCL INFO" gpurun_out/r2m/bench${N}_sharded.err | grep -i "error\|Traceback" -A8 | head -30
