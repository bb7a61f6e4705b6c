OUT=gpurun_out/r2r; mkdir -p $OUT
for t in fc00 fc10 fc11 fc21; do
  FSGM_LIB=$PWD/fsgm_b200/libfsgm_$t.so timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), d["stage_ms_per_step"], d["clocks"])
except Exception as e: print("$t parse failed", e)
PY
done
for t in fc11 fc21; do
  FSGM_LIB=$PWD/fsgm_b200/libfsgm_$t.so timeout 600 python -m pytest tests/test_epi_gpu.py -m gpu -x -q > $OUT/pytest_$t.log 2>&1; echo "pytest $t rc=$?"; tail -3 $OUT/pytest_$t.log
done
