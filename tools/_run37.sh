OUT=gpurun_out/r2P; mkdir -p $OUT
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 200 $OUT/bench.json; echo; tail -2 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"fused",d["e2e_fused"]["value"], "wall", d.get("bench_wall_s"), d["config"]["pairs_per_step_per_gpu"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"])
PY
