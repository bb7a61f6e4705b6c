timeout 900 python -m pytest tests/test_ng_gpu.py -m gpu -q -x -k "ng and not pydng" 2>&1 | tail -3
timeout 300 python tools/ng_occ.py
