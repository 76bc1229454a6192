#!/bin/bash
TAG=${1:-c4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { echo "== $*"; env "$@" timeout 120 python scripts/time_roi.py 2>&1 | sed 's/fwd general.*| bwd/bwd/' | tee -a $OUT/sweep.log; }
for d in 0 1 2 3 4 8 12 7 15; do run SOSWSOD_BWDQ_DEBUG=$d; done
