import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from fsgm_b200 import api
from oracle import pyramid_oracle as pyo
z = np.load("tests/golden/pyramid_a.npz")
ctx = api.Context(0)
L = int(z["numPyd"])
o = api.pyd_opts(numPyd=L, ver=int(z["ver"]), hor=int(z["hor"]))
mv, minC, lv = ctx.pyramidal_sgm(z["I0"], z["I1"], opts=o, levels=True)
for i in range(L - 1, -1, -1):
    d = lv[i] != z[f"mv_l{i}"]
    print("level", i, lv[i].shape, "mismatch", d.mean(), "x", d[0].mean(), "y", d[1].mean())
    if d.any():
        ys, xs = np.nonzero(d.any(0))
        print("  first", ys[:5], xs[:5], lv[i][:, ys[0], xs[0]], z[f"mv_l{i}"][:, ys[0], xs[0]])
# manual loop using the gateway
solver = lambda I1, I2, pre, rx, ry, agg, sub, P1, P2, diag, passes, adp: dict(zip(("bestD", "minC", "mvSub"), ctx.calc_pyd_cost_sgm(I1, I2, pre, rx, ry, agg, sub, P1, P2, diag, passes, adp)))
rmv, rmc, rlv = pyo.pyramidal_sgm(z["I0"], z["I1"], solver, numPyd=L, ver=int(z["ver"]), hor=int(z["hor"]))
for i in range(L - 1, -1, -1):
    print("manual level", i, (rlv[i] != z[f"mv_l{i}"]).mean())
