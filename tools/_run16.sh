OUT=gpurun_out/r2s; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"epi_cost_fused|hsweep|vsweep_kernel" -s 4 -c 4 -o $OUT/epiB -f python bench.py --skip A,C,D,strong_256 --no-cpu --steps 1 --warmup 1 --pairs 15 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu.log
ls -la $OUT
