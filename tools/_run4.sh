mkdir -p gpurun_out/r2l
N=16 timeout 300 python tools/pyd_quick.py > gpurun_out/r2l/pyd_plain.log 2>&1 && \
N=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pydl_sweep" -s 2 -c 1 -o gpurun_out/r2l/pydl4 python tools/pyd_quick.py > gpurun_out/r2l/pyd_ncu.log 2>&1; echo "ncu rc=$?"
