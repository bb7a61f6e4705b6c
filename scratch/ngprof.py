import sys; sys.path.insert(0,'.')
import numpy as np, torch
from fsgm_b200 import api, synth
ctx = api.Context(0); ctx.use_torch_stream()
W,hs,n=1242,48,148
sp = synth.flow_pair(W, hs, seed=2, umax=20, vmax=4)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
I1 = t(np.stack([sp["I1"]] * n)); I2 = t(np.stack([sp["I2"]] * n))
mC = torch.empty((n, hs, W), dtype=torch.int32, device="cuda"); fl = torch.empty((n, 2, hs, W), dtype=torch.float64, device="cuda")
f=lambda: ctx.calc_cost_sgm_ng_dev(I1, I2, 6, 32, mC, fl, seeds=list(range(1, n + 1)))
f(); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); f(); e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)
print('ng us per pixel per CTA', ms*1e3/(W*hs), ' kitti pairs/s (148 CTAs)', n*W*hs/(ms*1e-3)/(1242*375))
