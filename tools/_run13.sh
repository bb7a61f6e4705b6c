mkdir -p gpurun_out/r2o
N=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyd_cost_list -c 1 -s 1 -o gpurun_out/r2o/pcl -f python tools/pyd_quick.py > gpurun_out/r2o/ncu_pcl.log 2>&1; echo "rc=$?"
