OUT=gpurun_out/r2D; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_pyd_gpu.py tests/test_pyramid_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for t in main prev; do
  L=$PWD/fsgm_b200/libfsgm_$t.so; [ $t = main ] && L=$PWD/fsgm_b200/libfsgm.so
  FSGM_LIB=$L timeout 600 python bench.py --skip A,D,strong_256 --no-cpu --steps 5 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), {k: round(v["value"],1) for k,v in d["workloads"]["C"].items()})
    for k,v in d["workloads"]["C"].items(): print(k, {a: round(b,2) if isinstance(b,float) else b for a,b in v.get("stage_ms_per_step",{}).items()})
except Exception as e: print("$t parse failed", e)
PY
done
