OUT=gpurun_out/r2G; mkdir -p $OUT
timeout 900 python -m pytest tests/test_epi_gpu.py -m gpu -x -q -k "odd_vmax or large_finite or wild" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
