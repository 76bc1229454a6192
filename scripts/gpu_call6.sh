#!/bin/bash
TAG=${1:-c6}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "roi_pool_fast_backward or roi_pool_backward" 2>&1 | tail -3
run() { echo "== $*"; env "$@" timeout 120 python scripts/time_roi.py 2>&1 | sed 's/fwd general.*| bwd/bwd/' | tee -a $OUT/sweep.log; }
run SOSWSOD_BWDQ_P=0
run SOSWSOD_BWDQ_P=5
run SOSWSOD_BWDQ_P=6
run SOSWSOD_BWDQ_P=6 SOSWSOD_BWDQ_LQ=2
