import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from fsgm_b200 import api, synth
ctx = api.Context(0); ctx.use_torch_stream()
W,H,D=1242,375,256
p = synth.epipolar_pair(W,H,D,seed=1)
o = api.epi_opts(paths=8)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
hp = {k: pin(p[k]) for k in ("I1","I2","Pd0","dirn","O")}
for _ in range(3): ctx.calc_cost_sgm(hp["I1"],hp["I2"],D,0.3,hp["Pd0"],hp["dirn"],hp["O"],6,64,opts=o)
t0=time.perf_counter()
for _ in range(20): ctx.calc_cost_sgm(hp["I1"],hp["I2"],D,0.3,hp["Pd0"],hp["dirn"],hp["O"],6,64,opts=o)
print('single pair host gateway (pinned): %.2f ms' % ((time.perf_counter()-t0)/20*1e3))
for _ in range(3): ctx.calc_cost_sgm(p["I1"],p["I2"],D,0.3,p["Pd0"],p["dirn"],p["O"],6,64,opts=o)
t0=time.perf_counter()
for _ in range(20): ctx.calc_cost_sgm(p["I1"],p["I2"],D,0.3,p["Pd0"],p["dirn"],p["O"],6,64,opts=o)
print('single pair host gateway (pageable): %.2f ms' % ((time.perf_counter()-t0)/20*1e3))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).cuda()
d = {k: t(p[k]) for k in ("I1","I2","Pd0","dirn","O")}
b = torch.empty((1,H,W),dtype=torch.int32,device='cuda'); m=torch.empty_like(b)
f = lambda: ctx.calc_cost_sgm_dev(d["I1"],d["I2"],D,0.3,d["Pd0"],d["dirn"],d["O"],6,64,b,m,opts=o)
for _ in range(3): f()
torch.cuda.synchronize(); ctx.profile(True); ctx.profile_reset()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print('single pair device-resident: %.3f ms' % (e0.elapsed_time(e1)/20), {k:round(v[0]/20,3) for k,v in ctx.profile_read().items()})
