OUT=gpurun_out/r2K; mkdir -p $OUT
CASES=40 timeout 800 python tools/stress_pyd.py > $OUT/stress_pyd.log 2>&1; echo "stress rc=$?"; tail -5 $OUT/stress_pyd.log
