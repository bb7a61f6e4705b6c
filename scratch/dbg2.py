import sys; sys.path.insert(0,'.')
import numpy as np, torch, ctypes as C
from fsgm_b200 import synth, api
from oracle import pyoracle as po
ctx = api.Context(0); ctx.use_torch_stream()
W,H,D=75,70,64
p = synth.epipolar_pair(W, H, D, seed=D)
rng = np.random.default_rng(D)
O = p["O"].copy()
O[0, :4] = [np.nan, 1e300, -1e300, 3e9]
O[1, :5] = [-5.0, 2.0 ** 31, -(2.0 ** 31) - 7, np.inf, -np.inf]
O[H // 2, W // 2] = 2.0 ** 31 / 0.1
Pd0 = p["Pd0"] + rng.normal(0, 0.3, p["Pd0"].shape)
Pd0[0, 2, :6] = [0.5, 1.5, -2.5, 2.5, 3.5, W + 0.5]
Pd0[1, 3, :4] = [2147483648.5, 2147483649.0, -1e9, 0.49999999999999994 + 1]
cen1, cen2 = po.port_census(p["I1"]), po.port_census(p["I2"])
lib = po._port()
raw = np.empty((H, W, D), np.uint8); want = np.empty((H, W, D), np.uint8)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
lib.orc_epi_cost_raw(vp(cen1), vp(cen2), W, H, D, C.c_double(0.3), vp(Pd0), vp(p["dirn"]), vp(O), vp(raw))
lib.orc_box5(vp(raw), W, H, D, vp(want))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
Cv = torch.empty((1, H, W, D), dtype=torch.uint8, device="cuda")
ctx.epi_cost_dev(t(cen1.view(np.int32)[None]), t(cen2.view(np.int32)[None]), D, 0.3, t(Pd0[None]), t(p["dirn"][None]), t(O[None]), None, Cv)
g = Cv.cpu().numpy()[0]
bad = np.argwhere(g != want)
print(len(bad), bad[:20])
ys,xs = np.unique(bad[:,0]), np.unique(bad[:,1]); print(ys, xs)
