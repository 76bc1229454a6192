mkdir -p gpurun_out/r2g
N=$1
for mode in sharded allreduce; do
  NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $mode > gpurun_out/r2g/bench${N}_$mode.json 2> gpurun_out/r2g/bench${N}_$mode.err
  echo "== N=$N $mode rc=$?"; head -c 420 gpurun_out/r2g/bench${N}_$mode.json; echo; grep -v "NCCL INFO" gpurun_out/r2g/bench${N}_$mode.err | tail -4
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload detect --images 5000 > gpurun_out/r2g/detect${N}.json 2> gpurun_out/r2g/detect${N}.err
echo "== detect N=$N rc=$?"; head -c 500 gpurun_out/r2g/detect${N}.json; echo; tail -3 gpurun_out/r2g/detect${N}.err
