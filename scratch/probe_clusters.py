import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from fsgm_b200 import api
lib = api.lib()
lib.fsgm_debug_max_clusters.argtypes = [C.c_int, C.c_size_t, C.c_int]
def smem(D, Wk, ndir=3, warps=20):
    return 2*Wk*D + ndir*Wk*D + ((ndir*Wk*4 + 15) & ~15) + 4*(D+16) + warps*D*2 + 64
for cs in range(1, 17):
    Wk = (1242 + cs - 1)//cs
    s = smem(256, Wk)
    if s > 227*1024: print(cs, Wk, s, "too big"); continue
    for thr in (640,):
        k = lib.fsgm_debug_max_clusters(cs, s, thr)
        print("cs", cs, "Wk", Wk, "smem", s, "clusters", k, "SMs", k*cs)
