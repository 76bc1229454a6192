#!/bin/bash
TAG=${1:-b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python bench.py ${@:2} > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
b=json.load(open("$OUT/bench.json"))
print({k:b[k] for k in ["value","ms_per_step","gpu_launches","clocks"]}); print(b["e2e"]); print(b["roofline"]["achieved"], b["roofline"]["frac"], b["roofline"]["gemm_share_of_step"]); print(b["cpu_baseline"])
PY
