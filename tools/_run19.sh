OUT=gpurun_out/r2v; mkdir -p $OUT
for t in mb3 mb2; do
  FSGM_LIB=$PWD/fsgm_b200/libfsgm_$t.so timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), {k: round(v,2) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print("$t parse failed", e)
PY
done
