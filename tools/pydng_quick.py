#!/usr/bin/env python
"""Timing of calc_pyd_cost_sgm_ng at KITTI size (not a test: prints)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsgm_b200 import api, synth

ctx = api.Context(0); ctx.use_torch_stream()
W, H = 1242, 375
n = int(os.environ.get("N", "8"))
fp = synth.flow_pair(W, H, seed=2, umax=20, vmax=10)
I1 = torch.from_numpy(np.stack([fp["I1"]] * n)).cuda(); I2 = torch.from_numpy(np.stack([fp["I2"]] * n)).cuda()
rng = np.random.default_rng(1)
priors = {"zero": np.zeros((2, H, W)), "blocks": 2.0 * np.repeat(np.repeat(rng.integers(-3, 4, (2, (H + 7) // 8, (W + 7) // 8)), 8, 1), 8, 2)[:, :H, :W].astype(np.float64)}
for r in (1, 2):
    for name, pr in priors.items():
        mv = torch.from_numpy(np.stack([pr] * n)).cuda()
        mC = torch.empty((n, H, W), dtype=torch.int32, device="cuda"); fl = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
        ctx.calc_pyd_cost_sgm_ng_dev(I1, I2, mv, r, 5, 1, 6, 32, mC, fl); torch.cuda.synchronize()
        ctx.profile(True); ctx.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2): ctx.calc_pyd_cost_sgm_ng_dev(I1, I2, mv, r, 5, 1, 6, 32, mC, fl)
        e1.record(); torch.cuda.synchronize()
        st = ctx.profile_read(); ctx.profile(False)
        ms = e0.elapsed_time(e1) / 2
        print(f"r={r} prior={name} n={n}: {ms:.1f} ms per batch -> {n / ms * 1e3:.0f} pairs/s; stages:", {k: round(v[0] / 2, 2) for k, v in st.items()}, flush=True)
ctx.close()
