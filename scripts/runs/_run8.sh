# usage: _run8.sh N "modes" detect_images
mkdir -p gpurun_out/r2h
N=$1
for mode in $2; do
  NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $mode > gpurun_out/r2h/bench${N}_$mode.json 2> gpurun_out/r2h/bench${N}_$mode.err
  echo "== N=$N $mode rc=$?"; head -c 300 gpurun_out/r2h/bench${N}_$mode.json; echo; grep -v "NCCL INFO" gpurun_out/r2h/bench${N}_$mode.err | tail -4
done
if [ "$3" != "0" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload detect --images $3 > gpurun_out/r2h/detect${N}.json 2> gpurun_out/r2h/detect${N}.err
echo "== detect N=$N rc=$?"; head -c 700 gpurun_out/r2h/detect${N}.json; echo; tail -3 gpurun_out/r2h/detect${N}.err
fi
