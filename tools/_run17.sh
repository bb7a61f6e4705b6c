OUT=gpurun_out/r2t; mkdir -p $OUT
for t in 128 96 75 63 54 47 0; do
  timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 --tune-cost-rows $t > $OUT/bench_ty$t.json 2> $OUT/bench_ty$t.err; echo "ty$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_ty$t.json")); print("ty$t", round(d["value"],1), {k: round(v,2) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print("ty$t parse failed", e)
PY
done
timeout 600 python -m pytest tests/test_epi_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
