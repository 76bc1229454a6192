#!/bin/bash
OUT=gpurun_out/c7; mkdir -p $OUT
one() { tag=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline > $OUT/$tag.json 2> $OUT/$tag.err; python - <<PY
import json
b=json.load(open("$OUT/$tag.json")); r=b["roofline"]
print("$tag", round(b["ms_per_step"],3), "gemm", round(sum(d["avg_ms"] for d in r["detail"]),3), "roi", round(r["roi_pool"]["fwd"]["ms_per_step"]+r["roi_pool"]["bwd"]["ms_per_step"],3), "e2e", round(b["e2e"]["ms_per_step"],3), b["clocks"].get("samples"), b["clocks"].get("sm_mhz"))
PY
}
one smi100_a SOSWSOD_SMI_MS=100
one nosmi_a SOSWSOD_NO_SMI=1
one smi100_b SOSWSOD_SMI_MS=100
one nosmi_b SOSWSOD_NO_SMI=1
one smi500_a SOSWSOD_SMI_MS=500
one smi100_c SOSWSOD_SMI_MS=100
