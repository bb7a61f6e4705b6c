mkdir -p gpurun_out/r2h
N=16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2h/pyd_launches.csv python tools/pyd_quick.py > gpurun_out/r2h/pyd_ncu.log 2>&1; echo "pyd ncu rc=$?"
