# one-GPU evidence run of round 2: tests, bench line, launch list, ncu --set full of the top kernels, detect workload
OUT=gpurun_out/r2u; mkdir -p $OUT
python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > $OUT/tests.log 2>&1; tail -3 $OUT/tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; head -c 300 $OUT/bench.json; echo; tail -2 $OUT/bench.err
export SOSWSOD_PRE_WARMUP=0 SOSWSOD_SETTLE_BLOCKS=0 SOSWSOD_NO_SMI=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --blocks 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; tail -1 $OUT/ncu_launch_bench.log | head -c 200; echo
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|roi_pool_fwd_fast|roi_pool_bwd_q|sgd_multi" -s 42 -c 14 -o $OUT/prof_top python bench.py --steps 2 --warmup 3 --blocks 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; tail -2 $OUT/ncu_full.log | head -c 300; echo
unset SOSWSOD_PRE_WARMUP SOSWSOD_SETTLE_BLOCKS SOSWSOD_NO_SMI
timeout 900 python bench.py --workload detect --images 1500 > $OUT/detect1.json 2> $OUT/detect1.err; head -c 600 $OUT/detect1.json; echo; tail -2 $OUT/detect1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --shape coco > $OUT/bench_coco.json 2> $OUT/bench_coco.err; head -c 200 $OUT/bench_coco.json; echo; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --refine-k 4 --proposals 4000 > $OUT/bench_k4_r4000.json 2> $OUT/bench_k4_r4000.err; head -c 200 $OUT/bench_k4_r4000.json; echo
