OUT=gpurun_out/r2y; mkdir -p $OUT
timeout 900 python -m pytest tests/test_epi_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for t in main prev; do
  L=$PWD/fsgm_b200/libfsgm_$t.so; [ $t = main ] && L=$PWD/fsgm_b200/libfsgm.so
  FSGM_LIB=$L timeout 300 python bench.py --skip C,D,strong_256 --no-cpu --steps 10 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), {k: round(v,2) for k,v in d["stage_ms_per_step"].items()}, "A", round(d["workloads"]["A"]["value"],1))
except Exception as e: print("$t parse failed", e)
PY
done
