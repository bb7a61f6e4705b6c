mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_pyd_gpu.py tests/test_pyramid_gpu.py -m gpu -q -x > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2f/pytest.log
N=${NP:-32} timeout 300 python tools/pyd_quick.py > gpurun_out/r2f/pyd_plain.log 2>&1; cat gpurun_out/r2f/pyd_plain.log
