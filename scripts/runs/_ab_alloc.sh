OUT=gpurun_out/r2l; mkdir -p $OUT
for i in 1 2; do
  python -m pytest tests/test_gpu_engine.py -m gpu -q -p no:cacheprovider -k "not bench_shape" > /dev/null 2>&1
  for conf in default expandable; do
    if [ $conf = expandable ]; then export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True; else unset PYTORCH_CUDA_ALLOC_CONF; fi
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_${conf}_$i.json 2> $OUT/bench_${conf}_$i.err
    python - <<PY
import json
b=json.load(open("$OUT/bench_${conf}_$i.json")); r=b["roofline"]
big=[d for d in r["detail"] if d["m"]*d["n"]*d["k"]>5e11]
print("$conf $i", "ms/step", round(b["ms_per_step"],3), "e2e", round(b["e2e"]["ms_per_step"],3), "slow", r["slow_mode_seen_in_event_pass"], " | ".join(f"{'wgrad' if d['a_mn'] else ('dgrad' if d['b_mn'] else 'fwd')} avg {d['avg_ms']:.3f} max {d['max_ms']:.3f}" for d in big))
PY
  done
done
