OUT=gpurun_out/r2O; mkdir -p $OUT
for P in 60 90 120; do
  timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 --pairs $P > $OUT/bench_$P.json 2> $OUT/bench_$P.err; echo "$P rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$P.json")); print("pairs=$P", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fused", round(d["e2e_fused"]["value"],1), round(d["ms_per_step"],2))
except Exception as e: print("$P parse failed", e); print(open("$OUT/bench_$P.err").read()[-400:])
PY
done
