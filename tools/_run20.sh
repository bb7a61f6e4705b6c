OUT=gpurun_out/r2w; mkdir -p $OUT
for t in pipe1 pipe2 pipe3; do
  FSGM_LIB=$PWD/fsgm_b200/libfsgm_$t.so timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 > $OUT/bench_$t.json 2> $OUT/bench_$t.err; echo "$t rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$t.json")); print("$t", round(d["value"],1), {k: round(v,2) for k,v in d["stage_ms_per_step"].items()})
except Exception as e: print("$t parse failed", e)
PY
  FSGM_LIB=$PWD/fsgm_b200/libfsgm_$t.so timeout 600 python -m pytest tests/test_epi_gpu.py -m gpu -x -q -k "fused or stages or kitti" > $OUT/pytest_$t.log 2>&1; echo "pytest $t rc=$?"; tail -2 $OUT/pytest_$t.log
done
