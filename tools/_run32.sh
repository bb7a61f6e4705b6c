OUT=gpurun_out/r2J; mkdir -p $OUT
CASES=60 timeout 600 python tools/stress_cost.py > $OUT/stress_cost.log 2>&1; echo "stress rc=$?"; tail -4 $OUT/stress_cost.log
