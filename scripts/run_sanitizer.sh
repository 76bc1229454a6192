#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled synchronisation (VERDICT r01 "missing" #7).  Logs -> $1 (default
# gpurun_out/sanitizer); the summaries are committed under profiles/.
OUT=${1:-gpurun_out/sanitizer}
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  for what in roi gemm; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_targets.py $what > $OUT/${tool}_${what}.log 2>&1
    echo "== $tool $what rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' $OUT/${tool}_${what}.log | tr '\n' ';')"
  done
done
