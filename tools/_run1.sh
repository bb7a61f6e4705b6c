tools/gpu_round.sh r2d tests smoke bench ref ncu
N=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d/pyd_launches.csv python tools/pyd_quick.py > gpurun_out/r2d/pyd_ncu.log 2>&1; echo "pyd ncu rc=$?"
