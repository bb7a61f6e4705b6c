import sys, os, time, numpy as np, cv2
sys.path.insert(0, "/root/repo")
from oracle import pyoracle as po, pyramid_oracle as pyo
ex = "/root/reference/proj/example"
rgb2g = lambda im: np.floor(0.298936021293775*im[...,2] + 0.587043074451121*im[...,1] + 0.114020904255103*im[...,0] + 0.5).astype(np.uint8)
I0 = np.ascontiguousarray(rgb2g(cv2.imread(f"{ex}/000000_10.png", cv2.IMREAD_UNCHANGED).astype(np.float64)))
I1 = np.ascontiguousarray(rgb2g(cv2.imread(f"{ex}/000000_11.png", cv2.IMREAD_UNCHANGED).astype(np.float64)))
gt = cv2.imread(f"{ex}/000000_10_gtFlow.png", cv2.IMREAD_UNCHANGED)[..., ::-1].astype(np.float32)
t0 = time.time()
mv, mc, _ = pyo.pyramidal_sgm(I0, I1, lambda *a: po.ref_pyd(*a, stages=False), numPyd=5)
print("oracle seconds", time.time() - t0)
u, v = mv[0].astype(np.float32), mv[1].astype(np.float32)
gu, gv, val = (gt[..., 0] - 32768) / 64, (gt[..., 1] - 32768) / 64, gt[..., 2] > 0
e = np.hypot(u - gu, v - gv); mag = np.hypot(gu, gv)
print("outliers %.2f %%  epe %.3f" % (100 * ((e > 3) & (e > 0.05 * mag))[val].mean(), e[val].mean()))
