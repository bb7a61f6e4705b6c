mkdir -p gpurun_out/r2p
NG_H=48 NG_OCC=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ng_pipe -c 1 -s 1 -o gpurun_out/r2p/ng -f python tools/ng_prof.py > gpurun_out/r2p/ncu_ng.log 2>&1; echo "ng rc=$?"
N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pydng -c 4 -s 4 -o gpurun_out/r2p/pydng -f python tools/pydng_quick.py > gpurun_out/r2p/ncu_pydng.log 2>&1; echo "pydng rc=$?"
N=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pyd_cost|pydl|pyd_uniform" -c 6 -s 18 -o gpurun_out/r2p/pyd -f python tools/pyd_quick.py > gpurun_out/r2p/ncu_pyd.log 2>&1; echo "pyd rc=$?"
ls -la gpurun_out/r2p
