#!/bin/bash
mkdir -p gpurun_out/r2z
timeout 75 python -m pytest tests -x -q -m gpu > gpurun_out/r2z/tests_pitch.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2z/tests_pitch.log
timeout 25 python scripts/time_roi_fwd.py 2>&1 | tee gpurun_out/r2z/time_pitch.log
