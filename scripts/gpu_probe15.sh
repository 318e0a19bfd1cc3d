#!/bin/bash
python - <<'PY' 2>&1 | grep -v Warn | tail -n 14
import os, sys, torch
sys.path.insert(0, '.')
from torch.profiler import ProfilerActivity, profile
from coordx_b200 import modules as M
torch.manual_seed(0)
kw = dict(type='sine', in_features=2, out_features=1, hidden_features=256, mode='mlp', quiet=True)
unsplit = M.SingleBVPNet(split_mlp=False, **kw).cuda()
pts = torch.rand(1, 2048 * 2048, 2, device='cuda') * 2 - 1
with torch.no_grad():
    for _ in range(2):
        unsplit.net(pts)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        unsplit.net(pts)
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
for k, t, n in rows[:10]:
    print(f'{t:10.1f} us {n:5d}x  {k[:100]}')
PY
