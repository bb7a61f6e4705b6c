NG_H=24 NG_OCC=1 timeout 300 python tools/ng_prof.py 2>&1 | tail -19 | sort
