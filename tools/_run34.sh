OUT=gpurun_out/r2L; mkdir -p $OUT
timeout 900 python -m pytest tests/test_epi_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for t in main prev; do
  L=$PWD/fsgm_b200/libfsgm_$t.so; [ $t = main ] && L=$PWD/fsgm_b200/libfsgm.so
  for P in 32 47; do
  FSGM_LIB=$L timeout 300 python bench.py --skip A,C,D,strong_256 --no-cpu --steps 10 --warmup 3 --pairs $P > $OUT/bench_${t}_$P.json 2> $OUT/bench_${t}_$P.err; echo "$t $P rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${t}_$P.json")); print("$t pairs=$P", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["ms_per_step"],2))
except Exception as e: print("$t parse failed", e)
PY
  done
done
