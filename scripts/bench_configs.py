#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration on ONE B200 (bench.py measures only the headline one).

For each named shape: grid points/s of the two-stage inference (CUDA-graph replay of forward_split_channel per
axis + forward_split_fusion, train_img.py:125-135) and iterations/s of the eager training step
(model(dict) -> loss -> backward -> Adam, training.py:112-146), CUDA-event timed after warm-up, L2 flushed between
forward steps.  FLOP figures are SURVEY.md 8(d)'s algorithmic ones (the per-point Hd x Hd GEMM is credited even
where the separable sum path does not execute it).  Prints one JSON line per configuration.

    python scripts/bench_configs.py [--only NAME] [--steps 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coordx_b200 import modules as M            # noqa: E402
from coordx_b200.runtime import CapturedGrid, CapturedTrainStep    # noqa: E402

SUSTAINED_TF = 1406.8


def configs():
    vid = dict(type='sine', in_features=3, out_features=3, hidden_features=1024, split_rule=[1, 2], fusion_operator='sum')
    img = lambda op: dict(type='sine', in_features=2, out_features=1, hidden_features=256, fusion_operator=op)
    occ = lambda op: dict(type='sine', in_features=3, out_features=1, hidden_features=256, fusion_operator=op)
    return [
        # name, ctor kwargs, inference grid, training grid
        ('cfg1_img256_sum', img('sum'), (256, 256), (256, 256)),
        ('cfg1_img512_sum', img('sum'), (512, 512), (512, 512)),
        ('cfg2_img2048_prod', img('prod'), (2048, 2048), (2048, 2048)),
        ('cfg2_img2048_sum', img('sum'), (2048, 2048), (2048, 2048)),
        ('cfg3_video_st_sum_h1024', vid, (300, 512 * 512), (150, 1992)),
        ('cfg4_occ256_prod', occ('prod'), (256, 256, 256), (32, 32, 32)),
        ('cfg4_occ256_sum', occ('sum'), (256, 256, 256), (32, 32, 32)),
        ('cfg5_img4k_prod', img('prod'), (2160, 3840), (2160, 3840)),
        ('cfg5_occ512_prod_slab64', occ('prod'), (64, 512, 512), None),      # one GPU's slab of 512^3 over 8 GPUs
    ]


def flops_fwd(kw, shape, f):
    Hd, O = kw['hidden_features'], kw['out_features']
    P = 1
    for n in shape:
        P *= n
    fl = P * (2 * Hd * Hd + 2 * Hd * O)
    for n, fc in zip(shape, f):
        fl += n * (2 * fc * Hd + 2 * 2 * Hd * Hd)
    return fl, P


def axes_for(shape, f):
    out = []
    for n, fc in zip(shape, f):
        if fc == 1:
            out.append(torch.linspace(-1, 1, n, device='cuda')[:, None].contiguous())
        else:
            out.append(torch.rand(n, fc, device='cuda') * 2 - 1)
    return out


def timed(fn, steps, warm, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--precision', default='auto')
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for name, kw, gshape, tshape in configs():
        if args.only and args.only not in name:
            continue
        torch.manual_seed(0)
        model = M.SingleBVPNet(split_mlp=True, quiet=True, mode='mlp', precision=args.precision, **kw).cuda()
        f = model.net.feat_per_channel
        rec = {'config': name, 'hidden': kw['hidden_features'], 'fusion': kw['fusion_operator'], 'grid': list(gshape),
               'precision': args.precision}
        # ---- inference
        axes = axes_for(gshape, f)
        cg = CapturedGrid(model, axes)
        med, best = timed(cg.replay, args.steps, 5, flush)
        fl, P = flops_fwd(kw, gshape, f)
        rec.update(fwd_ms=round(med, 4), fwd_ms_min=round(best, 4), grid_points_per_s=P / med * 1e3,
                   fwd_tflops_algorithmic=fl / med * 1e-9, fwd_frac_sustained=fl / med * 1e-9 / SUSTAINED_TF)
        del cg
        # ---- training step
        if tshape is not None:
            C = len(tshape)
            coords = []
            for c, a in enumerate(axes_for(tshape, f)):
                sh = [1] * (C + 1) + [a.shape[-1]]
                sh[c + 1] = a.shape[0]
                coords.append(a.reshape(sh))
            Pt = 1
            for n in tshape:
                Pt *= n
            gt = torch.rand(1, Pt, kw['out_features'], device='cuda') * 2 - 1
            opt = torch.optim.Adam(model.parameters(), lr=1e-4)

            def step():
                out = model({'coords_split': coords})['model_out']
                loss = ((out - gt) ** 2).mean()
                opt.zero_grad(set_to_none=False)
                loss.backward()
                opt.step()

            med, best = timed(step, args.steps, 5)
            flt, _ = flops_fwd(kw, tshape, f)
            rec.update(train_grid=list(tshape), train_ms=round(med, 4), train_it_per_s=1e3 / med,
                       train_tflops_algorithmic=3 * flt / med * 1e-9, train_frac_sustained=3 * flt / med * 1e-9 / SUSTAINED_TF)
            # the same iteration replayed from one CUDA graph (runtime.CapturedTrainStep)
            cap = CapturedTrainStep(model, coords, gt, loss='mse', lr=1e-4)
            med, best = timed(cap.replay, args.steps, 5)
            rec.update(train_graph_ms=round(med, 4), train_graph_it_per_s=1e3 / med,
                       train_graph_frac_sustained=3 * flt / med * 1e-9 / SUSTAINED_TF)
            del cap
        print(json.dumps(rec), flush=True)
        del model
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
