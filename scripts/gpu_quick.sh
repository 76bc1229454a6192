#!/bin/bash
# quick GPU check: selected tests + ROI timing.  gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag> "<pytest -k expr>"'
TAG=${1:-quick}; K=${2:-roi_pool}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "$K" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest.log
timeout 300 python scripts/time_roi.py 2>&1 | tee $OUT/time_roi.log
