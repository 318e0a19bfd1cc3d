#!/usr/bin/env python
"""Split vs un-split on one B200 (the comparison the reference's speed-up is defined against, train_img.py:115-135):
the 2048 x 2048 image evaluated (a) by the split MLP on the grid, (b) by the un-split FCBlock on the 4.19 M points
through the same kernels (modules._FCAsSplit), (c) by the un-split FCBlock as plain torch ops (the reference's path
on this GPU), (d) by the split MLP on the same 4.19 M SCATTERED points.  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coordx_b200 import modules as M            # noqa: E402
from coordx_b200.runtime import CapturedGrid    # noqa: E402


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    n = 2048
    torch.manual_seed(0)
    kw = dict(type='sine', in_features=2, out_features=1, hidden_features=256, mode='mlp', quiet=True)
    split = M.SingleBVPNet(split_mlp=True, fusion_operator='prod', **kw).cuda()
    unsplit = M.SingleBVPNet(split_mlp=False, **kw).cuda()
    line = torch.linspace(-1, 1, n, device='cuda')
    pts = torch.stack(torch.meshgrid(line, line, indexing='ij'), -1).reshape(1, -1, 2).contiguous()
    P = n * n
    rec = {'grid': [n, n], 'points': P}
    with torch.no_grad():
        cg = CapturedGrid(split, [line[:, None].contiguous(), line[:, None].contiguous()])
        rec['split_grid_ms'] = timed(cg.replay)
        rec['split_scattered_ms'] = timed(lambda: split.net(pts))
        rec['unsplit_kernels_ms'] = timed(lambda: unsplit.net(pts))
        unsplit.net.fused_impl = 'eager'
        rec['unsplit_torch_eager_ms'] = timed(lambda: unsplit.net(pts), steps=5, warm=2)
    for k in list(rec):
        if k.endswith('_ms'):
            rec[k[:-3] + '_Gpts_per_s'] = P / rec[k] * 1e-6
    rec['note'] = ('un-split = 3 hidden 256x256 layers per point (2 of them as fp32-accurate 3xTF32 GEMMs, the last + output layer in the '
                   'bf16 tcgen05 trunk); split on the grid = 1 per point')
    print(json.dumps(rec))


if __name__ == '__main__':
    main()
