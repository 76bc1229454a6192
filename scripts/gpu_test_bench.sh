#!/bin/bash
# full GPU parity suite + default bench line.  gpurun --timeout 1500 -- 'bash scripts/gpu_test_bench.sh <tag>'
TAG=${1:-tb}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
b=json.load(open("$OUT/bench.json"))
print({k:b[k] for k in ["value","ms_per_step","gpu_launches","clocks"]}); print(b["e2e"]); print(b["roofline"]["achieved"], b["roofline"]["frac"], b["roofline"]["gemm_share_of_step"]); print(b["cpu_baseline"])
PY
