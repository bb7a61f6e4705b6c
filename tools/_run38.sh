OUT=gpurun_out/r2R; mkdir -p $OUT
for t in 2 3 6; do
  timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 12 --warmup 3 --tune-stage-waves $t > $OUT/bench_sw$t.json 2> $OUT/bench_sw$t.err; echo "sw$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_sw$t.json")); print("stage_waves=$t value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fused", round(d["e2e_fused"]["value"],1), "lat", round(d["latency_ms_single_pair"]["host_gateway"],3))
except Exception as e: print("parse failed", e); print(open("$OUT/bench_sw$t.err").read()[-300:])
PY
done
