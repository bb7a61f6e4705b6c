OUT=gpurun_out/r2C; mkdir -p $OUT
N=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pydl_sweep|pydl_wta|pyd_cost_px" -s 6 -c 3 -o $OUT/pyd -f python tools/pyd_quick.py > $OUT/ncu_pyd.log 2>&1; echo "pyd rc=$?"; tail -3 $OUT/ncu_pyd.log
ls -la $OUT
