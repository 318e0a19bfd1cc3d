#!/usr/bin/env python
"""Pipeline timeline of the forward trunk kernel (CTA 0): per iteration, when the MMA issuer, the two
epilogue groups and producer warp 8 waited / worked, in SM clocks.  Run on the GPU box."""
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
x = torch.linspace(-1, 1, H, device='cuda')[:, None]; y = torch.linspace(-1, 1, W, device='cuda')[:, None]
tr = torch.zeros(4 * 96 * 4, dtype=torch.int64, device='cuda')
with torch.no_grad():
    fx = m.forward_split_channel(x, 0).unsqueeze(1); fy = m.forward_split_channel(y, 1).unsqueeze(0)
    for _ in range(3):
        m.forward_split_fusion([fx, fy])
    lib.coordx_profile_set_trace(tr.data_ptr())
    m.forward_split_fusion([fx, fy])
    torch.cuda.synchronize()
    lib.coordx_profile_set_trace(None)
t = tr.cpu().numpy().reshape(4, 96, 4)
t0 = t[0, 0, 0]
np.save(os.path.join('gpurun_out', 'fwd_trace.npy'), t)
print('it |  MMA: wait_tempty wait_full issue | period || EPI(grp 0): wait_tfull work || PROD w8: wait_empty work')
for it in range(8, 40):
    mma = t[0, it]; e = t[1, it]; pr = t[3, it]
    period = t[0, it + 1, 0] - mma[0]
    print(f'{it:3d} | {mma[1]-mma[0]:6d} {mma[2]-mma[1]:6d} {mma[3]-mma[2]:6d} | {period:6d} || {e[1]-e[0]:6d} {e[2]-e[1]:6d} (tfull@{e[1]-mma[3]:6d} after issue) || {pr[1]-pr[0]:6d} {pr[2]-pr[1]:6d}')
per = np.diff(t[0, 8:90, 0])
print('epi0 post (exchange+store)', (t[1, 8:90, 3] - t[1, 8:90, 2]).mean(), 'epi1 post', (t[2, 8:90, 3] - t[2, 8:90, 2]).mean(), 'epi0 loop gap', (t[1, 9:90, 0] - t[1, 8:89, 3]).mean())
print('mean period', per.mean(), 'epi work mean', (t[1:3, 8:90, 2] - t[1:3, 8:90, 1])[(t[1:3, 8:90, 2] > 0)].mean())
