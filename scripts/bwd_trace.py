#!/usr/bin/env python
"""Pipeline timeline of the dH backward kernel (CTA 0), SM clocks.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the stamps are a build option: scripts/build_trace_lib.sh
os.environ.setdefault('COORDX_B200_LIB', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'coordx_b200', '_build', 'libcoordx_trace.so'))
import numpy as np
import torch
import __graft_entry__ as G
G.build()
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
PASS = int(os.environ.get('TRACE_PASS', '2'))      # 0 = dW pass, 2 = dH pass (last axis); pass 1 streams stored tiles (no stamps)
lib.coordx_profile_set_bwd_trace(tr.data_ptr(), PASS)
step()
torch.cuda.synchronize()
lib.coordx_profile_set_bwd_trace(None, 1)
t = tr.cpu().numpy().reshape(4, 96, 4)
np.save(os.path.join('gpurun_out', 'bwd_trace.npy'), t)
print('t | MMA(G2): wait_dz wait_dhempty issue | period || E1: wait_z wait_dzempty(after b0 start) work_total || E2: wait_dh work || PROD: wait_empty work stage')
for it in range(8, 36):
    mm = t[0, it]; e1 = t[1, it]; e2 = t[2, it]; pr = t[3, it]
    print(f'{it:3d} | {mm[1]-mm[0]:6d} {mm[2]-mm[1]:6d} {mm[3]-mm[2]:6d} | {t[0,it+1,0]-mm[0]:6d} || {e1[1]-e1[0]:6d} {e1[2]-e1[1]:6d} {e1[3]-e1[1]:6d} || {e2[1]-e2[0]:6d} {e2[2]-e2[1]:6d} || {pr[1]-pr[0]:6d} {pr[2]-pr[1]:6d} {pr[3]-pr[2]:6d}')
print('mean period', np.diff(t[0, 8:90, 0]).mean())
