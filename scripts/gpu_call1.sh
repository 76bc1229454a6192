#!/bin/bash
# queued ROI backward A/B + TTA path + full suite + bench.  gpurun --timeout 1500 -- 'bash scripts/gpu_call1.sh <tag>'
TAG=${1:-c1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q -k "roi_pool_fast_backward or roi_pool_backward" > $OUT/pytest_roi.log 2>&1; RC=$?
echo "roi bwd pytest exit $RC"; tail -5 $OUT/pytest_roi.log
if [ $RC -ne 0 ]; then export SOSWSOD_ROI_BWD=turn; echo "FALLING BACK to turn-token backward for the rest"; fi
timeout 200 python scripts/time_roi.py 2>&1 | tee $OUT/time_roi_default.log
SOSWSOD_ROI_BWD=turn timeout 200 python scripts/time_roi.py 2>&1 | tee $OUT/time_roi_turn.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cut -c1-400 $OUT/bench.json
timeout 300 python scripts/bench_detect.py > $OUT/bench_detect.json 2> $OUT/bench_detect.err; echo "detect exit $?"; cat $OUT/bench_detect.json; tail -3 $OUT/bench_detect.err
timeout 300 python scripts/bench_detect.py --scales 480 576 672 768 864 960 1056 1152 --refine-k 4 > $OUT/bench_detect_16v.json 2>> $OUT/bench_detect.err; cat $OUT/bench_detect_16v.json
