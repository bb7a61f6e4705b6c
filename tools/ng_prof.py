#!/usr/bin/env python
"""One calc_cost_sgm_ng call on 148 x OCC strips of 1242 x H (for ncu): NG_H rows, NG_OCC resident pairs per SM."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fsgm_b200 import api, synth

ctx = api.Context(0); ctx.use_torch_stream()
W, H, occ = 1242, int(os.environ.get("NG_H", "8")), int(os.environ.get("NG_OCC", "3"))
n = torch.cuda.get_device_properties(0).multi_processor_count * occ
fp = synth.flow_pair(W, H, seed=2, umax=20, vmax=4)
I1 = torch.from_numpy(np.stack([fp["I1"]] * n)).cuda(); I2 = torch.from_numpy(np.stack([fp["I2"]] * n)).cuda()
mC = torch.empty((n, H, W), dtype=torch.int32, device="cuda"); fl = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
ctx.tune(3, occ)
for _ in range(2):
    ctx.calc_cost_sgm_ng_dev(I1, I2, 6, 32, mC, fl, seeds=list(range(1, n + 1))); torch.cuda.synchronize()
ctx.close()
