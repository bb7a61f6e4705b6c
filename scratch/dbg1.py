import sys; sys.path.insert(0,'.')
import numpy as np, torch
from fsgm_b200 import synth, api
from oracle import pyoracle as po
ctx = api.Context(0); ctx.use_torch_stream()
W,H,D=96,64,32
p = synth.epipolar_pair(W,H,D,seed=W+H+D)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for first in (4,8,4):
    ref = po.ref_epi(p['I1'],p['I2'],D,p['vMax'],p['Pd0'],p['dirn'],p['O'],6,64,paths=first)
    prt = po.port_epi(p['I1'],p['I2'],D,p['vMax'],p['Pd0'],p['dirn'],p['O'],6,64,paths=first)
    Sp = torch.empty((1,H,W,D),dtype=torch.int16,device='cuda'); b = torch.empty((1,H,W),dtype=torch.int32,device='cuda'); m=torch.empty_like(b)
    ctx.epi_aggregate_dev(t(ref['C'][None]), t(p['I1'][None]), 6,64, t(p['O'][None]), p['vMax'], b, m, Sp=Sp, opts=api.epi_opts(paths=first))
    g = Sp.cpu().numpy().view(np.uint16)[0].astype(np.uint32)
    print(first, 'gpu', g[0,0,:4], 'ref', ref['Sp'][0,0,:4], 'port', prt['Sp'][0,0,:4])
