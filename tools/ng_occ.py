#!/usr/bin/env python
"""A/B of the resident-pairs knob of the calc_cost_sgm_ng kernel (fsgm_tune key 3) on 1242x48 strips: pixels/s per setting."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsgm_b200 import api, synth

ctx = api.Context(0); ctx.use_torch_stream()
W, H = 1242, 48
sm = torch.cuda.get_device_properties(0).multi_processor_count
fp = synth.flow_pair(W, H, seed=2, umax=20, vmax=4)
for occ in (1, 2, 3):
    n = sm * occ
    I1 = torch.from_numpy(np.stack([fp["I1"]] * n)).cuda(); I2 = torch.from_numpy(np.stack([fp["I2"]] * n)).cuda()
    mC = torch.empty((n, H, W), dtype=torch.int32, device="cuda"); fl = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
    ctx.tune(3, occ)
    seeds = list(range(1, n + 1))
    ctx.calc_cost_sgm_ng_dev(I1, I2, 6, 32, mC, fl, seeds=seeds); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.calc_cost_sgm_ng_dev(I1, I2, 6, 32, mC, fl, seeds=seeds); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"occ {occ}: {n} strips in {ms:.1f} ms -> {n * W * H / ms / 1e3:.2f} Mpx/s = {n * W * H / ms * 1e3 / (1242 * 375):.1f} KITTI pairs/s", flush=True)
ctx.close()
