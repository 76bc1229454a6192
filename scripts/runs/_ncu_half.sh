#!/bin/bash
mkdir -p gpurun_out/r2z
timeout 140 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd_half -c 4 -f -o gpurun_out/r2z/half_fwd python scripts/runs/_ncu_half.py > gpurun_out/r2z/ncu_half.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2z/ncu_half.log; ls -la gpurun_out/r2z/*.ncu-rep
