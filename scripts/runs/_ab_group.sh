OUT=gpurun_out/r2j; mkdir -p $OUT
python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/tests.log 2>&1; tail -2 $OUT/tests.log
for g in 8 16 32 8 16; do
  SOSWSOD_GEMM_GROUP_M=$g timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_g$g.json 2> $OUT/bench_g$g.err
  python - <<PY
import json
b=json.load(open("$OUT/bench_g$g.json")); r=b["roofline"]
print("group_m=$g", "ms/step", round(b["ms_per_step"],3), [round(x,2) for x in b["blocks_ms_per_step"]], "gemm ms", round(r["gemm_ms_per_step"],3), "TF", round(r["achieved"]), " ".join(f"{d['avg_ms']:.3f}" for d in r["detail"]), "clk", b["clocks"]["sm_mhz"])
PY
done
