#!/usr/bin/env python
"""Batch producer of the video path at BASELINE configs[2]'s shape: 300 x 512 x 512 RGB video,
~298 844-point batches (z_num frames x M pixels).  Times (a) the reference's formulation - CPU gather + pinned H2D copy
of the batch, what Implicit3DWrapper.__getitem__ + the DataLoader do (oracle port, run on this box's host cores), and
(b) coordx_b200.dataio.Implicit3DWrapper: device-resident data, one gather kernel; with the reference's host-side index
draws and with device-side draws.  Prints one JSON line.  Run on the GPU box."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coordx_b200 import dataio as G      # noqa: E402
from oracle import dataio_np as D        # noqa: E402  (CPU baseline leg only)


class Video:
    def __init__(self, vid):
        self.vid, self.shape, self.channels = vid, vid.shape[:-1], vid.shape[-1]

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        return self.vid


T, H, W, C = 300, 512, 512, 3
frac = 298844 / (T * H * W)
vid = torch.rand(T, H, W, C, generator=torch.Generator().manual_seed(0)).numpy()
ds = G.Implicit3DWrapper(Video(vid), sidelength=(T, H, W), split_coord=True, sample_fraction=frac)
np.random.seed(0)
torch.manual_seed(0)


def timed_gpu(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


ms_host_draw = timed_gpu(lambda: ds.sample(), 200)
ms_dev_draw = timed_gpu(lambda: ds.sample(device_rng=True), 200)
# kernel alone
idz = torch.randperm(T, device='cuda')[:150].contiguous()
idxy = torch.randint(0, H * W, (1992,), device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    ds._gather(idz, idxy)
e0.record()
for _ in range(200):
    ds._gather(idz, idxy)
e1.record()
torch.cuda.synchronize()
ms_kernel = e0.elapsed_time(e1) / 200
rows = 150 * 1992

# reference formulation on the host: gather from the CPU array + pinned copy to the GPU
data_cpu = torch.from_numpy(((vid - 0.5) / 0.5).reshape(-1, C))
pin = torch.empty(400000, C).pin_memory()
dev = torch.empty(400000, C, device='cuda')


def ref_step():
    z_num = np.random.randint(round(T * 0.1), round(T * 0.9))
    cz = torch.randperm(T)[:z_num]
    cxy = torch.randint(0, H * W, (round(T * H * W * frac / z_num),))
    cid = (cz[:, None] * (H * W) + cxy[None, :]).reshape(-1)
    b = data_cpu[cid, :]
    pin[:b.shape[0]].copy_(b)
    dev[:b.shape[0]].copy_(pin[:b.shape[0]], non_blocking=True)
    torch.cuda.synchronize()


for _ in range(3):
    ref_step()
t0 = time.perf_counter()
for _ in range(30):
    ref_step()
ms_ref = (time.perf_counter() - t0) / 30 * 1e3
print(json.dumps({'workload': 'video batch producer, 300x512x512 RGB, ~298844-point batches (z_num x M), dataio.py:575-583',
                  'reference_cpu_gather_plus_h2d_ms': ms_ref, 'host_cores': os.cpu_count(),
                  'gpu_ms_per_batch_reference_draws': ms_host_draw, 'gpu_ms_per_batch_device_draws': ms_dev_draw,
                  'gather_kernel_ms': ms_kernel, 'gather_rows': rows,
                  'gather_kernel_GBps': rows * C * 4 * 2 / (ms_kernel * 1e-3) / 1e9,
                  'note': 'gather kernel: 12-byte rows scattered over 943 MB: every row costs a 32-byte sector read; '
                          'algorithmic bytes = rows x C x 4 read + written'}))
