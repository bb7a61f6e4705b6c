OUT=gpurun_out/r2Q; mkdir -p $OUT

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench2 rc=$?"; head -c 600 $OUT/bench_2gpu.json; echo; python - <<PY
import json
d=json.load(open("$OUT/bench_2gpu.json"))
print({k: (v.get("value") if isinstance(v,dict) else v) for k,v in d.get("workloads",{}).items()})
print(d.get("e2e",{}).get("value"), d.get("e2e_fused",{}).get("value"))
PY
tail -3 $OUT/bench_2gpu.err
