mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_pyd_gpu.py tests/test_pyramid_gpu.py -m gpu -q -x > gpurun_out/r2g/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2g/pytest.log
N=32 timeout 300 python tools/pyd_quick.py > gpurun_out/r2g/pyd_plain32.log 2>&1; cat gpurun_out/r2g/pyd_plain32.log
N=16 timeout 300 python tools/pyd_quick.py > gpurun_out/r2g/pyd_plain.log 2>&1 && \
N=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pydl_sweep|pyd_cost_sep" -s 4 -c 2 -o gpurun_out/r2g/pydl2 python tools/pyd_quick.py > gpurun_out/r2g/pyd_ncu.log 2>&1; echo "ncu rc=$?"
