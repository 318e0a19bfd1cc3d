#!/usr/bin/env python
"""Pipeline timeline of the feature-gradient pass (trunk_bwd_dfeat_kernel, CTA 0), SM clocks.  Run on the GPU box."""
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
lib.coordx_profile_set_bwd_trace(tr.data_ptr(), 1)
step()
torch.cuda.synchronize()
lib.coordx_profile_set_bwd_trace(None, 1)
t = tr.cpu().numpy().reshape(4, 96, 4)
base = t[0, 40, 0]
print('MMA per sub-tile s: top | wait_dz wait_accempty issue || E2 g0 / g1 per tile: top | wait_accfull read(->accempty) rest || loader per sub-tile: top | wait_dzempty')
for s_ in range(40, 90):
    m0 = t[0, s_]; e0 = t[1, s_ // 2]; e1 = t[2, s_ // 2]; ld = t[3, s_]
    print(f'{s_:3d} | {m0[0]-base:7d} {m0[1]-m0[0]:5d} {m0[2]-m0[1]:5d} {m0[3]-m0[2]:5d} || {e0[0]-base:7d} {e0[1]-e0[0]:5d} {e0[2]-e0[1]:5d} {e0[3]-e0[2]:5d} | '
          f'{e1[0]-base:7d} {e1[1]-e1[0]:5d} {e1[2]-e1[1]:5d} {e1[3]-e1[2]:5d} || {ld[0]-base:7d} {ld[1]-ld[0]:5d}')
print('MMA period per sub-tile', np.diff(t[0, 20:90, 0]).mean(), 'wait_dz', (t[0, 20:90, 1] - t[0, 20:90, 0]).mean(), 'wait_accempty', (t[0, 20:90, 2] - t[0, 20:90, 1]).mean(),
      'issue', (t[0, 20:90, 3] - t[0, 20:90, 2]).mean())
for g in (1, 2):
    e = t[g, 10:45]
    print(f'E2 g{g-1}: period per tile', np.diff(e[:, 0]).mean(), 'wait_accfull', (e[:, 1] - e[:, 0]).mean(), 'read', (e[:, 2] - e[:, 1]).mean(), 'rest', (e[:, 3] - e[:, 2]).mean())
