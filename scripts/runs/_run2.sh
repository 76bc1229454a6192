mkdir -p gpurun_out/r2e
for mode in sharded allreduce; do
  NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $mode > gpurun_out/r2e/bench2_$mode.json 2> gpurun_out/r2e/bench2_$mode.err
  echo "== $mode rc=$?"; tail -c 1200 gpurun_out/r2e/bench2_$mode.json; grep -c "NCCL INFO" gpurun_out/r2e/bench2_$mode.err; grep -i "nranks\|NVLS" gpurun_out/r2e/bench2_$mode.err | head -4; grep -v "NCCL INFO" gpurun_out/r2e/bench2_$mode.err | tail -8
done
