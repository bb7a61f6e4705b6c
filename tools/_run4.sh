mkdir -p gpurun_out/r2g
N=16 timeout 300 python tools/pyd_quick.py > gpurun_out/r2g/pyd_plain.log 2>&1 && \
N=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pydl_sweep|pydl_wta" -s 4 -c 2 -o gpurun_out/r2g/pydl python tools/pyd_quick.py > gpurun_out/r2g/pyd_ncu.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r2g/pyd_plain.log
