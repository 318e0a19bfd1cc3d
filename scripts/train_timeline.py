#!/usr/bin/env python
"""Kernel timeline (start, duration, stream) of ONE replay of the captured training step (runtime.CapturedTrainStep)
of a named configuration: shows what is on the critical path besides the trunk kernels.
   python scripts/train_timeline.py cfg2_img2048_prod"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as B          # noqa: E402
from coordx_b200 import modules as M   # noqa: E402
from coordx_b200.runtime import CapturedTrainStep   # noqa: E402

name = sys.argv[1]
_, kw, gshape, tshape = [c for c in B.configs() if c[0] == name][0]
torch.manual_seed(0)
model = M.SingleBVPNet(split_mlp=True, quiet=True, mode='mlp', **kw).cuda()
f = model.net.feat_per_channel
C = len(tshape)
coords = []
for c, a in enumerate(B.axes_for(tshape, f)):
    sh = [1] * (C + 1) + [a.shape[-1]]
    sh[c + 1] = a.shape[0]
    coords.append(a.reshape(sh))
P = 1
for n in tshape:
    P *= n
gt = torch.rand(1, P, kw['out_features'], device='cuda') * 2 - 1
cap = CapturedTrainStep(model, coords, gt, loss='mse', lr=1e-4)
for _ in range(5):
    cap.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    cap.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type.name == 'CUDA']
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
end = max(e.time_range.end for e in evs)
print(f'# {name}: one graph replay, {len(evs)} device activities, span {(end - t0):.1f} us')
print('#  start_us   dur_us  stream  kernel')
for e in evs:
    print(f'{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f}  {getattr(e, "stream", getattr(e, "device_index", 0))!s:>6}  {e.name[:100]}')
