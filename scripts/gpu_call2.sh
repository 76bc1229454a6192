#!/bin/bash
# queued ROI backward tuning: poll back-off / prep warps / depth, + ncu of both backward kernels
TAG=${1:-c2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { echo "== $*"; env "$@" timeout 120 python scripts/time_roi.py 2>&1 | sed 's/fwd general.*| bwd/bwd/' | tee -a $OUT/sweep.log; }
run SOSWSOD_BWDQ_PREP_SLEEP=0
run SOSWSOD_BWDQ_PREP_SLEEP=100
run SOSWSOD_BWDQ_PREP_SLEEP=200
run SOSWSOD_BWDQ_PREP_SLEEP=500
run SOSWSOD_BWDQ_PREP_SLEEP=200 SOSWSOD_BWDQ_ACC_SLEEP=50
run SOSWSOD_BWDQ_PREP_SLEEP=200 SOSWSOD_BWDQ_P=3
run SOSWSOD_BWDQ_PREP_SLEEP=200 SOSWSOD_BWDQ_P=2
run SOSWSOD_BWDQ_PREP_SLEEP=200 SOSWSOD_BWDQ_P=1
run SOSWSOD_BWDQ_PREP_SLEEP=200 SOSWSOD_BWDQ_P=4 SOSWSOD_BWDQ_D=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_q -s 4 -c 1 -f -o $OUT/bwd_q python scripts/time_roi.py > $OUT/ncu_q.log 2>&1; echo "ncu q exit $?"
SOSWSOD_ROI_BWD=turn timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_fast -s 4 -c 1 -f -o $OUT/bwd_turn python scripts/time_roi.py > $OUT/ncu_turn.log 2>&1; echo "ncu turn exit $?"
timeout 300 python -m pytest tests/test_gpu_tta.py -m gpu -x -q 2>&1 | tail -3
ls -la $OUT
