#!/usr/bin/env python
"""Quick check + timing of the pyd cluster path against the scanline kernels (not a test: prints)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsgm_b200 import api, synth

ctx = api.Context(0); ctx.use_torch_stream()
small = len(sys.argv) > 1 and sys.argv[1] == "small"
if small:
    fp = synth.flow_pair(60, 24, seed=3, umax=3, vmax=2)
    mv = np.round(np.random.default_rng(1).normal(0, 1.5, (2, 24, 60)))
    for cs in (1, 2, 4):
        ctx.tune(5, cs)
        a = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0)
        ctx.tune(5, -1)
        b = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0)
        print("cs", cs, "equal:", [bool(np.array_equal(x, y, equal_nan=True)) for x, y in zip(a, b)], "diff px:", int((a[0] != b[0]).sum()), flush=True)
    ctx.tune(5, 0); ctx.close(); sys.exit(0)
W, H = 1242, 375
n = int(os.environ.get("N", "24"))
fps = [synth.flow_pair(W, H, seed=1 + i, umax=20, vmax=10) for i in range(4)]
I0 = torch.from_numpy(np.stack([fps[i % 4]["I1"] for i in range(n)])).cuda(); I1 = torch.from_numpy(np.stack([fps[i % 4]["I2"] for i in range(n)])).cuda()
mv = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda"); mC = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
mv2 = torch.empty_like(mv); mC2 = torch.empty_like(mC)
for r in (5, 4):
    o = api.pyd_opts(numPyd=3, ver=r, hor=r)
    for mode in (0, -1):
        ctx.tune(5, mode)
        out = (mv, mC) if mode == 0 else (mv2, mC2)
        ctx.pyramidal_sgm_dev(I0, I1, out[0], out[1], opts=o); torch.cuda.synchronize()
        ctx.profile(True); ctx.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): ctx.pyramidal_sgm_dev(I0, I1, out[0], out[1], opts=o)
        e1.record(); torch.cuda.synchronize()
        st = ctx.profile_read(); ctx.profile(False)
        ms = e0.elapsed_time(e1) / 3
        print(f"r={r} mode={'lane' if mode == 0 else 'scanline'} n={n}: {ms:.2f} ms per batch -> {n / ms * 1e3:.0f} pairs/s; stages ms/batch:",
              {k: round(v[0] / 3, 2) for k, v in st.items()}, flush=True)
    print("  lane == scanline:", bool(torch.equal(mv, mv2) and torch.equal(mC, mC2)), flush=True)
ctx.tune(5, 0); ctx.close()
