import sys; sys.path.insert(0,'.')
import ctypes as C, torch
from fsgm_b200 import api
torch.cuda.init(); torch.zeros(1).cuda()
l = api.lib()
for cs in (1,2,4,8,16):
    for smem in (100*1024, 120*1024, 160*1024, 219*1024, 226*1024):
        for th in (512, 1024):
            print(cs, smem//1024, th, l.fsgm_debug_max_clusters(cs, C.c_size_t(smem), th))
