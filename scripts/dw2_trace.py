#!/usr/bin/env python
"""Pipeline timeline of the dW pass (trunk_bwd_dw2_kernel, CTA 0), SM clocks per 64-point sub-tile.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the stamps are a build option: scripts/build_trace_lib.sh
os.environ.setdefault('COORDX_B200_LIB', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'coordx_b200', '_build', 'libcoordx_trace.so'))
import numpy as np
import torch
from coordx_b200 import modules as M, _lib
lib = _lib.load()
torch.manual_seed(0)
H = W = 2048
m = M.SingleBVPNet(type='sine', out_features=1, hidden_features=256, split_mlp=True, fusion_operator='prod', quiet=True, precision='bf16').cuda()
xs = torch.linspace(-1, 1, H, device='cuda').reshape(1, H, 1, 1); ys = torch.linspace(-1, 1, W, device='cuda').reshape(1, 1, W, 1)
gt = torch.rand(1, H * W, 1, device='cuda') * 2 - 1
tr = torch.zeros(4 * 96 * 4, dtype=torch.int64, device='cuda')
def step():
    m.zero_grad()
    out = m({'coords_split': [xs, ys]})['model_out']
    ((out - gt) ** 2).mean().backward()
for _ in range(2):
    step()
lib.coordx_profile_set_bwd_trace(tr.data_ptr(), 0)
step()
torch.cuda.synchronize()
lib.coordx_profile_set_bwd_trace(None, 1)
t = tr.cpu().numpy().reshape(4, 96, 4)
np.save(os.path.join('gpurun_out', 'dw2_trace.npy'), t)
base = t[1, 20, 0]
print('s | E1 pq=0: top wait compute tail period | E1 pq=1: top wait compute tail || MMA: top g1_issue wait_dz g3 || PROD(t = s/2): top wait_empty write dout')
for s_ in range(20, 56):
    a = t[1, s_]; b = t[2, s_]; m0 = t[0, s_]; pr = t[3, s_ // 2]
    print(f'{s_:3d} | {a[0]-base:7d} {a[1]-a[0]:5d} {a[2]-a[1]:5d} {a[3]-a[2]:5d} {t[1,s_+1,0]-a[0]:5d} | {b[0]-base:7d} {b[1]-b[0]:5d} {b[2]-b[1]:5d} {b[3]-b[2]:5d} || '
          f'{m0[0]-base:7d} {m0[1]-m0[0]:5d} {m0[2]-m0[1]:5d} {m0[3]-m0[2]:5d} || {pr[0]-base:7d} {pr[1]-pr[0]:5d} {pr[2]-pr[1]:5d} {pr[3]-pr[2]:5d}')
ev = t[1, 20:90]
print('pq 0: mean period per sub-tile', np.diff(ev[:, 0]).mean(), ' wait', (ev[:, 1] - ev[:, 0]).mean(), ' compute', (ev[:, 2] - ev[:, 1]).mean(),
      ' tail', (ev[:, 3] - ev[:, 2]).mean())
