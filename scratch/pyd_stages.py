import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from fsgm_b200 import api, synth
ctx = api.Context(0); ctx.use_torch_stream()
W, H, n = 1242, 375, 8
fp = synth.flow_pair(W, H, seed=1, umax=20, vmax=10)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
I1 = t(np.stack([fp["I1"]] * n)); I2 = t(np.stack([fp["I2"]] * n))
mv = torch.zeros((n, 2, H, W), dtype=torch.float64, device="cuda")
bD = torch.empty((n, H, W), dtype=torch.int32, device="cuda"); mC = torch.empty_like(bD)
ms_ = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
for r in (5, 4):
    for _ in range(2): ctx.calc_pyd_cost_sgm_dev(I1, I2, mv, r, r, 2, 1, 6, 32, 1, 2, 0, bD, mC, ms_)
    torch.cuda.synchronize(); ctx.profile(True); ctx.profile_reset()
    for _ in range(3): ctx.calc_pyd_cost_sgm_dev(I1, I2, mv, r, r, 2, 1, 6, 32, 1, 2, 0, bD, mC, ms_)
    torch.cuda.synchronize()
    print("r", r, {k: round(v[0] / 3 / n, 3) for k, v in ctx.profile_read().items()}, "ms per pair")
    ctx.profile(False)
