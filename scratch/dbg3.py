import sys; sys.path.insert(0,'.')
import numpy as np, torch, ctypes as C
from fsgm_b200 import synth, api
from oracle import pyoracle as po
ctx = api.Context(0); ctx.use_torch_stream()
W,H,D=75,70,64
p = synth.epipolar_pair(W, H, D, seed=D)
cen1, cen2 = po.port_census(p["I1"]), po.port_census(p["I2"])
lib = po._port()
vp = lambda a: a.ctypes.data_as(C.c_void_p)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for name,val in [('nan',np.nan),('1e300',1e300),('-1e300',-1e300),('inf',np.inf),('-inf',-np.inf),('3e9',3e9),('-5',-5.0),('1e18',1e18),('-1e18',-1e18),('1e12',1e12)]:
    O = p["O"].copy(); O[5,7]=val
    raw = np.empty((H, W, D), np.uint8); want = np.empty((H, W, D), np.uint8)
    lib.orc_epi_cost_raw(vp(cen1), vp(cen2), W, H, D, C.c_double(0.3), vp(p["Pd0"]), vp(p["dirn"]), vp(O), vp(raw))
    lib.orc_box5(vp(raw), W, H, D, vp(want))
    Cv = torch.empty((1, H, W, D), dtype=torch.uint8, device="cuda")
    ctx.epi_cost_dev(t(cen1.view(np.int32)[None]), t(cen2.view(np.int32)[None]), D, 0.3, t(p["Pd0"][None]), t(p["dirn"][None]), t(O[None]), None, Cv)
    g = Cv.cpu().numpy()[0]
    bad = np.argwhere(g != want)
    print(name, len(bad), np.unique(bad[:,2])[:10] if len(bad) else '')
