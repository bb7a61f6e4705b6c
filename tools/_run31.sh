OUT=gpurun_out/r2I; mkdir -p $OUT
N=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pydl_sweep" -s 2 -c 1 -o $OUT/pydl32 -f python tools/pyd_quick.py > $OUT/ncu_pyd.log 2>&1; echo "pyd rc=$?"; tail -3 $OUT/ncu_pyd.log
