#!/bin/bash
# ncu --set full of the ROI kernels at bench shape.  gpurun --timeout 900 -- 'bash scripts/gpu_ncu_roi.sh <tag> <kernel-regex> <skip> <count>'
TAG=${1:-ncu_roi}; KRE=${2:-roi_pool}; SKIP=${3:-0}; CNT=${4:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o $OUT/prof \
   python scripts/time_roi.py > $OUT/ncu.log 2>&1; echo "ncu exit $?"; tail -5 $OUT/ncu.log
