#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py tests/test_gpu_tta.py -x -q -m gpu -k "roi or tta" > gpurun_out/r2y/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2y/tests.log
timeout 300 python scripts/time_roi_fwd.py 2>&1 | tee gpurun_out/r2y/time_half2.log
