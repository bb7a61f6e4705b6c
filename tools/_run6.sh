N=32 R=5 timeout 300 python tools/pyd_quick.py 2>&1 | grep lane
