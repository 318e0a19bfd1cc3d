#!/usr/bin/env python
"""BASELINE.json configs[4]: the 4K (3840x2160) image and the 512^3 occupancy grid SHARDED over N B200s along axis 0
(coordx_b200/sharding.py) - strong scaling: the global grid is fixed, every rank owns a contiguous slab.

    python scripts/bench_sharded.py                                        # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_sharded.py

Inference: CUDA-graph replay of the two-stage API on the rank's slab (no collective).  Training (4K image): the eager
iteration model(dict) -> MSE -> backward -> ONE flat NCCL all-reduce of the weight gradients -> Adam.  Every number is
CUDA-event time on the device per step (median step), barrier + synchronize on both sides, MAX over ranks; rank 0 prints one JSON line per
workload."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coordx_b200 import modules as M, sharding                 # noqa: E402
from coordx_b200 import loss_functions as LF                   # noqa: E402
from coordx_b200.runtime import CapturedGrid, CapturedTrainStep   # noqa: E402

SUSTAINED_TF = 1406.8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--no-graph', action='store_true')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        per = sorted(a.elapsed_time(b) for a, b in evs)
        ms = per[len(per) // 2]                      # median step: the eager NCCL loop shows an occasional one-off ~180 ms stall
        if per[-1] > 3 * ms and rank == 0:
            print('# uneven steps (ms): ' + ' '.join(f'{v:.2f}' for v in per), file=sys.stderr, flush=True)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    workloads = [
        ('img4k_3840x2160_prod', dict(type='sine', in_features=2, out_features=1, hidden_features=256, fusion_operator='prod'),
         (2160, 3840), True),
        ('occ512_prod', dict(type='sine', in_features=3, out_features=1, hidden_features=256, fusion_operator='prod'),
         (512, 512, 512), False),
    ]
    for name, kw, shape, do_train in workloads:
        torch.manual_seed(0)                                   # identical replicas on every rank
        model = M.SingleBVPNet(split_mlp=True, quiet=True, mode='mlp', precision='bf16', **kw).to(dev)
        C = len(shape)
        a0, b0 = sharding.slab_range(shape[0], rank, world)
        axes = [torch.linspace(-1, 1, n, device=dev)[:, None].contiguous() for n in shape]
        my_axes = [axes[0][a0:b0].contiguous()] + axes[1:]
        P = 1
        for n in shape:
            P *= n
        F = P * (2 * 256 * 256 + 2 * 256) + sum(n * (2 * 256 + 4 * 256 * 256) for n in shape)
        with torch.no_grad():
            cg = CapturedGrid(model, my_axes)
            ms = timed(cg.replay, args.steps, args.warmup)
            del cg
        rec = dict(workload=name, grid=list(shape), n_gpus=world, slab_rows=[a0, b0], sharding='axis 0, contiguous slabs',
                   fwd_ms=ms, grid_points_per_s=P / ms * 1e3, fwd_frac_of_sustained_peak_all_gpus=F / ms * 1e-9 / (SUSTAINED_TF * world))
        if do_train:
            coords = []
            for c, a in enumerate(my_axes):
                sh = [1] * (C + 1) + [1]
                sh[c + 1] = a.shape[0]
                coords.append(a.reshape(sh))
            gt = torch.rand(1, P, 1, generator=torch.Generator().manual_seed(0)) * 2 - 1
            gt = sharding.shard_target(gt, shape, rank, world).to(dev)
            opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

            def step():
                out = model({'coords_split': coords})['model_out']
                loss = LF.image_mse(None, {'model_out': out}, {'img': gt})['img_loss']
                opt.zero_grad(set_to_none=False)
                loss.backward()
                sharding.allreduce_grads(model.parameters(), world)
                opt.step()

            tms = timed(step, max(args.steps // 2, 5), 3)
            if not args.no_graph:
                # the same iteration (incl. the NCCL all-reduce) replayed from one CUDA graph per rank
                cap = CapturedTrainStep(model, coords, gt, loss='mse', lr=1e-4, world=world)
                gms = timed(cap.replay, max(args.steps // 2, 5), 3)
                rec.update(train_graph_ms=gms, train_graph_it_per_s=1e3 / gms,
                           train_graph_frac_of_sustained_peak_all_gpus=3 * F / gms * 1e-9 / (SUSTAINED_TF * world))
                del cap
            rec.update(train_ms=tms, train_it_per_s=1e3 / tms, train_frac_of_sustained_peak_all_gpus=3 * F / tms * 1e-9 / (SUSTAINED_TF * world),
                       train_note='eager iteration, one flat NCCL all-reduce of 198 657 gradient floats per step; slabs of 2160 rows are equal only when N divides 2160')
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del model
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
