OUT=gpurun_out/r2S; mkdir -p $OUT
timeout 600 python -m pytest tests/test_epi_gpu.py tests/test_geometry_gpu.py tests/test_proj_gpu.py tests/test_mex_stubs.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
