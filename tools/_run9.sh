timeout 900 python -m pytest tests/test_ng_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/pydng_quick.py
