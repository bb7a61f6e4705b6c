OUT=gpurun_out/r2H; mkdir -p $OUT
for t in main pl1_11 pl2_5 pl1_10; do
  L=$PWD/fsgm_b200/libfsgm_$t.so; [ $t = main ] && L=$PWD/fsgm_b200/libfsgm.so
  FSGM_LIB=$L timeout 600 python bench.py --skip A,D,strong_256 --no-cpu --steps 5 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), {k: round(v["value"],1) for k,v in d["workloads"]["C"].items()}, {k: round(v["stage_ms_per_step"]["pyd_sweep"],2) for k,v in d["workloads"]["C"].items()})
except Exception as e: print("$t parse failed", e)
PY
done
