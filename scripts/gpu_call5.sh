#!/bin/bash
TAG=${1:-c5}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "roi_pool_fast_backward or roi_pool_backward" 2>&1 | tail -3
run() { echo "== $*"; env "$@" timeout 120 python scripts/time_roi.py 2>&1 | sed 's/fwd general.*| bwd/bwd/' | tee -a $OUT/sweep.log; }
run SOSWSOD_BWDQ_P=0
run SOSWSOD_BWDQ_P=3
run SOSWSOD_BWDQ_P=2
run SOSWSOD_BWDQ_LQ=2
run SOSWSOD_BWDQ_LQ=1
run SOSWSOD_BWDQ_P=2 SOSWSOD_BWDQ_LQ=2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_q -s 4 -c 1 -f -o $OUT/bwd_q python scripts/time_roi.py > $OUT/ncu_q.log 2>&1; echo "ncu q exit $?"
