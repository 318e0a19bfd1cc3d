#!/usr/bin/env python
"""Per-kernel time of one eager forward / training step of a named configuration (torch.profiler, CUPTI).
   python scripts/kernel_breakdown.py cfg2_img2048_sum [train|fwd]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as B          # noqa: E402
from coordx_b200 import modules as M   # noqa: E402

name = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else 'train'
cfg = [c for c in B.configs() if c[0] == name][0]
_, kw, gshape, tshape = cfg
shape = tshape if mode == 'train' else gshape
torch.manual_seed(0)
model = M.SingleBVPNet(split_mlp=True, quiet=True, mode='mlp', **kw).cuda()
f = model.net.feat_per_channel
C = len(shape)
coords = []
for c, a in enumerate(B.axes_for(shape, f)):
    sh = [1] * (C + 1) + [a.shape[-1]]
    sh[c + 1] = a.shape[0]
    coords.append(a.reshape(sh))
P = 1
for n in shape:
    P *= n
gt = torch.rand(1, P, kw['out_features'], device='cuda') * 2 - 1
opt = torch.optim.Adam(model.parameters(), lr=1e-4)


def step():
    if mode == 'train':
        out = model({'coords_split': coords})['model_out']
        loss = ((out - gt) ** 2).mean()
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
    else:
        with torch.no_grad():
            model({'coords_split': coords})


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3, e.count / 3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f'{name} {mode}: {tot:.1f} us of kernels per step, {sum(r[2] for r in rows):.0f} launches')
for k, t, n in rows[:25]:
    print(f'{t:9.1f} us {n:5.1f}x  {k[:110]}')
