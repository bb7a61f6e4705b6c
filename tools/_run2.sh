mkdir -p gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e/pytest.log
N=16 timeout 300 python tools/pyd_quick.py > gpurun_out/r2e/pyd_plain.log 2>&1 && \
N=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pydv_kernel -s 4 -c 2 -o gpurun_out/r2e/pydv python tools/pyd_quick.py > gpurun_out/r2e/pyd_ncu.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r2e/pyd_plain.log
