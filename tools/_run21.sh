OUT=gpurun_out/r2x; mkdir -p $OUT
for t in main d0 d1p2 d1p4; do
  L=$PWD/fsgm_b200/libfsgm_$t.so; [ $t = main ] && L=$PWD/fsgm_b200/libfsgm.so
  FSGM_LIB=$L timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), {k: round(v,2) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print("$t parse failed", e)
PY
done
timeout 900 python -m pytest tests/test_epi_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
