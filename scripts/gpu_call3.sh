#!/bin/bash
# ROI backward: L2 prefetch distance x kernel variant x queue shape
TAG=${1:-c3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "roi_pool_fast_backward or roi_pool_backward" 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 120 python scripts/time_roi.py 2>&1 | sed 's/fwd general.*| bwd/bwd/' | tee -a $OUT/sweep.log; }
run SOSWSOD_BWD_PREFETCH=0
run SOSWSOD_BWD_PREFETCH=3
run SOSWSOD_BWD_PREFETCH=6
run SOSWSOD_BWD_PREFETCH=12
run SOSWSOD_BWD_PREFETCH=6 SOSWSOD_BWDQ_P=2 SOSWSOD_BWDQ_D=2
run SOSWSOD_BWD_PREFETCH=6 SOSWSOD_BWDQ_P=2 SOSWSOD_BWDQ_D=1
run SOSWSOD_BWD_PREFETCH=6 SOSWSOD_BWDQ_P=4 SOSWSOD_BWDQ_D=1
run SOSWSOD_BWD_PREFETCH=6 SOSWSOD_BWDQ_RT=16
run SOSWSOD_ROI_BWD=turn SOSWSOD_BWD_PREFETCH=0
run SOSWSOD_ROI_BWD=turn SOSWSOD_BWD_PREFETCH=4
run SOSWSOD_ROI_BWD=turn SOSWSOD_BWD_PREFETCH=8
