#!/bin/bash
# One GPU-box call: parity tests, bench line, ncu launch list, ncu --set full of the three dominant kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json | cut -c1-600
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"
export SOSWSOD_PRE_WARMUP=0 SOSWSOD_SETTLE_BLOCKS=0   # under ncu every launch is replayed: only the caller's warm-up steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:gemm_bf16_kernel|roi_pool_fwd|roi_pool_bwd" -s 18 -c 16 -f -o $OUT/prof_top \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
unset SOSWSOD_PRE_WARMUP SOSWSOD_SETTLE_BLOCKS
ls -la $OUT
timeout 300 python scripts/bench_detect.py > $OUT/bench_detect.json 2> $OUT/bench_detect.err; echo "detect exit $?"; cat $OUT/bench_detect.json
