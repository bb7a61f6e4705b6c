#!/usr/bin/env python
"""Randomised shape sweep of the fused epipolar cost kernel against the unfused pair of kernels (epi_raw_cost_kernel + epi_box5_kernel,
independent code pinned by the parity tests) and, for the smaller shapes, against the CPU oracle.  Not a test: prints and exits non-zero
on the first mismatch."""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsgm_b200 import api, synth
from oracle import pyoracle as po

ctx = api.Context(0); ctx.use_torch_stream()          # kernels on torch's stream: torch.equal below is ordered behind them
rng = np.random.default_rng(int(os.environ.get("SEED", "7")))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
lib = po._port()
vp = lambda a: a.ctypes.data_as(C.c_void_p)
n_cases = int(os.environ.get("CASES", "40"))
for it in range(n_cases):
    D = int(rng.choice([64, 128, 256]))
    W = int(rng.integers(5, 300)); H = int(rng.integers(3, 200))
    n = int(rng.integers(1, 4))
    vMax = float(rng.choice([0.3, 0.1, 0.6]))
    ps = [synth.epipolar_pair(W, H, D, seed=int(rng.integers(1 << 30))) for _ in range(n)]
    for p in ps:                                              # a few wild pixels per pair
        k = int(rng.integers(0, 4))
        ys, xs = rng.integers(0, H, k), rng.integers(0, W, k)
        p["O"][ys, xs] = rng.choice([1e7, -3e8, 2e9, np.nan, np.inf, 0.0], size=k)
    st = lambda key: np.stack([p[key] for p in ps])
    cen1 = np.stack([po.port_census(p["I1"]) for p in ps]); cen2 = np.stack([po.port_census(p["I2"]) for p in ps])
    a = torch.empty((n, H, W, D), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a); raw = torch.empty_like(a)
    args = (t(cen1.view(np.int32)), t(cen2.view(np.int32)), D, vMax, t(st("Pd0")), t(st("dirn")), t(st("O")))
    ctx.epi_cost_dev(*args, None, a)
    ctx.epi_cost_dev(*args, raw, b)
    ok = torch.equal(a, b)
    ora = ""
    if ok and W * H * D < 3_000_000:
        for i, p in enumerate(ps):
            r = np.empty((H, W, D), np.uint8); want = np.empty((H, W, D), np.uint8)
            lib.orc_epi_cost_raw(vp(cen1[i]), vp(cen2[i]), W, H, D, C.c_double(vMax), vp(p["Pd0"]), vp(p["dirn"]), vp(p["O"]), vp(r))
            lib.orc_box5(vp(r), W, H, D, vp(want))
            ok = ok and np.array_equal(a[i].cpu().numpy(), want)
        ora = " +oracle"
    print(f"case {it}: n={n} {W}x{H} D={D} vMax={vMax}: {'ok' if ok else 'MISMATCH'}{ora}", flush=True)
    if not ok:
        an, bn = a.cpu().numpy(), b.cpu().numpy()
        for i, p in enumerate(ps):
            r = np.empty((H, W, D), np.uint8); want = np.empty((H, W, D), np.uint8)
            lib.orc_epi_cost_raw(vp(cen1[i]), vp(cen2[i]), W, H, D, C.c_double(vMax), vp(p["Pd0"]), vp(p["dirn"]), vp(p["O"]), vp(r))
            lib.orc_box5(vp(r), W, H, D, vp(want))
            df, du = np.argwhere(an[i] != want), np.argwhere(bn[i] != want)
            print(f"  pair {i}: fused != oracle at {len(df)} voxels, unfused != oracle at {len(du)} voxels; raw(unfused) != oracle raw: {int((raw[i].cpu().numpy() != r).sum())}")
            if len(df):
                print("   fused diffs (y, x, d) head:", df[:6].tolist(), "ys", sorted(set(df[:, 0].tolist()))[:12], "xs", sorted(set(df[:, 1].tolist()))[:12])
                y, x, d = df[0]
                print("   values fused/unfused/oracle:", int(an[i][y, x, d]), int(bn[i][y, x, d]), int(want[y, x, d]))
                wy, wx = np.argwhere(~np.isfinite(p["O"]) | (np.abs(p["O"]) > 1e6)).T if True else (None, None)
                print("   wild pixels (y, x, O):", [(int(yy), int(xx), float(p["O"][yy, xx])) for yy, xx in zip(wy, wx)])
        sys.exit(1)
print("all cases agree")
