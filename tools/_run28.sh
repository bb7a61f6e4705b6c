OUT=gpurun_out/r2M; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err; echo "bench8 rc=$?"; head -c 400 $OUT/bench_8gpu.json; echo
python - <<PY
import json
d=json.load(open("$OUT/bench_8gpu.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "fused", d["e2e_fused"]["value"])
for k,v in d.get("workloads",{}).items(): print(k, {a:b for a,b in v.items() if a in ("value","unit","speedup_vs_single_gpu","bit_equal_to_single_gpu_call","single_gpu_ms_same_box")})
print(d.get("host_h2d_gbs",{}).get("per_rank"))
PY
tail -2 $OUT/bench_8gpu.err
