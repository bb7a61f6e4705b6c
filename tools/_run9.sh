timeout 900 python -m pytest tests/test_pyd_gpu.py tests/test_pyramid_gpu.py -m gpu -q -x 2>&1 | tail -3
N=16 timeout 300 python tools/pyd_quick.py
