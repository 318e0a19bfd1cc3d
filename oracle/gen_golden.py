"""Generate golden fixtures by RUNNING THE REFERENCE ITSELF (test infrastructure).

Runs only in the build container, where /root/reference exists.  Imports the
reference's modules.py / loss_functions.py UNCHANGED (through the container-only
`torchmeta` stand-in in oracle/_standin) on CPU/fp32, and writes small .npz
fixtures to tests/golden/.  The fixtures travel to the GPU box; this script and
/root/reference do not need to.

    python oracle/gen_golden.py            # (re)writes tests/golden/*.npz
    python oracle/gen_golden.py NAME ...   # only the named cases (seeds depend on the position in CASES)

Each fixture holds: cfg (json), the reference state_dict, per-axis coords, gt,
model_out of SingleBVPNet.forward({'coords_split': ...}) (modules.py:447-454),
branch features (ret_feat=True), forward_split_channel/forward_split_fusion
output (modules.py:463-474), the scattered-path output (modules.py:257-309),
the loss (loss_functions.py:8-12,56-60), every param.grad after one
loss.backward() (training.py:137-138) and the parameters after 3 Adam steps
(training.py:17,146; occupancy cases with clip_grad_norm_ 1.0, training.py:140-144).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_standin'))
sys.path.insert(0, '/root/reference')

# ImageDownsampling.__init__ (modules.py:488) calls .cuda() whenever sidelength is
# given; make that a no-op so the reference constructs on this GPU-less box.
torch.Tensor.cuda = lambda self, *a, **k: self

import modules            # noqa: E402  (the reference, unchanged)
import loss_functions     # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def axes_for(shape, feats, seed, mode):
    g = torch.Generator().manual_seed(seed)
    axes = []
    C = len(shape)
    for c, (n, f) in enumerate(zip(shape, feats)):
        if mode == 'linspace' and f == 1:
            a = torch.linspace(-1, 1, n)[:, None]
        else:
            a = torch.rand(n, f, generator=g) * 2 - 1
        sh = [1] * (C + 1) + [f]
        sh[c + 1] = n
        axes.append(a.reshape(sh).contiguous())
    return axes


CASES = [
    # name, ctor kwargs, grid shape, loss, coords mode
    ('img_sine_sum_h64', dict(type='sine', out_features=1, hidden_features=64, fusion_operator='sum'), (24, 40), 'mse', 'linspace'),
    ('img_sine_prod_h64', dict(type='sine', out_features=1, hidden_features=64, fusion_operator='prod'), (33, 17), 'mse', 'linspace'),
    ('img_relu_prod_h64', dict(type='relu', out_features=1, hidden_features=64, fusion_operator='prod'), (20, 31), 'mse', 'rand'),
    ('img_nerf_sum_h64', dict(type='relu', mode='nerf', out_features=1, hidden_features=64, fusion_operator='sum', sidelength=(64, 64)), (16, 24), 'mse', 'linspace'),
    ('img_rgb_sine_prod_h64', dict(type='sine', out_features=3, hidden_features=64, fusion_operator='prod'), (12, 20), 'mse', 'linspace'),
    ('video_st_sine_sum_h64', dict(type='sine', in_features=3, out_features=3, hidden_features=64, split_rule=[1, 2], fusion_operator='sum'), (7, 50), 'mse', 'rand'),
    ('video_txy_sine_prod_h64', dict(type='sine', in_features=3, out_features=3, hidden_features=64, split_rule=[1, 1, 1], fusion_operator='prod'), (5, 6, 7), 'mse', 'rand'),
    ('occ_sine_prod_h64', dict(type='sine', in_features=3, out_features=1, hidden_features=64, fusion_operator='prod'), (9, 10, 11), 'bce', 'rand'),
    ('occ_relu_sum_h64', dict(type='relu', in_features=3, out_features=1, hidden_features=64, fusion_operator='sum'), (8, 5, 13), 'bce', 'rand'),
    ('occ_nerf_prod_h64', dict(type='relu', mode='nerf', in_features=3, out_features=1, hidden_features=64, fusion_operator='prod'), (6, 7, 8), 'bce', 'rand'),
    ('occ_fourier_sum_h64', dict(type='relu', mode='nerf', in_features=3, out_features=1, hidden_features=64, fusion_operator='sum', freq_params=[6., 16], include_coord=False), (6, 7, 8), 'bce', 'rand'),
    ('img_al0_h32', dict(type='sine', out_features=1, hidden_features=32, approx_layers=0, fusion_operator='prod'), (9, 14), 'mse', 'linspace'),
    ('img_al1_h64', dict(type='sine', out_features=1, hidden_features=64, approx_layers=1, fusion_operator='prod'), (10, 12), 'mse', 'linspace'),
    ('img_al3_h64', dict(type='sine', out_features=1, hidden_features=64, approx_layers=3, fusion_operator='sum'), (10, 12), 'mse', 'linspace'),
    ('img_al4_h64', dict(type='sine', out_features=1, hidden_features=64, approx_layers=4, fusion_operator='prod'), (10, 12), 'mse', 'linspace'),
    ('img_al4_llf16_rgb', dict(type='relu', out_features=3, hidden_features=64, approx_layers=4, last_layer_features=16, fusion_operator='prod'), (10, 12), 'mse', 'linspace'),
    ('img_fba_h64', dict(type='sine', out_features=1, hidden_features=64, fusion_before_act=True, fusion_operator='sum'), (10, 12), 'mse', 'linspace'),
    ('img_fs2_h32', dict(type='sine', out_features=1, hidden_features=32, fusion_size=2, fusion_operator='prod'), (10, 12), 'mse', 'linspace'),
    ('img_fs2_reduced_h32', dict(type='relu', out_features=1, hidden_features=32, fusion_size=2, reduced=True, fusion_operator='sum'), (10, 12), 'mse', 'linspace'),
    ('img_code_scale2_h64', dict(type='sine', out_features=1, hidden_features=64, learn_code=True, act_scale=2, fusion_operator='prod'), (10, 12), 'mse', 'linspace'),
    ('img_nhl2_h64', dict(type='sine', out_features=1, hidden_features=64, num_hidden_layers=2, approx_layers=1, fusion_operator='sum'), (10, 12), 'mse', 'linspace'),
    # tensor-core sized (Hd=256): the shapes of BASELINE configs at reduced grid size
    ('img_sine_prod_h256', dict(type='sine', out_features=1, hidden_features=256, fusion_operator='prod'), (48, 200), 'mse', 'linspace'),
    ('img_sine_sum_h256', dict(type='sine', out_features=1, hidden_features=256, fusion_operator='sum'), (40, 130), 'mse', 'linspace'),
    ('occ_sine_prod_h256', dict(type='sine', in_features=3, out_features=1, hidden_features=256, fusion_operator='prod'), (6, 10, 150), 'bce', 'rand'),
    ('img_relu_prod_h256', dict(type='relu', out_features=1, hidden_features=256, fusion_operator='prod'), (30, 128), 'mse', 'rand'),
    # sum-fused sine nets with three axes / a 2-D axis and RGB output (the separable trunk path)
    ('occ_sine_sum_h64', dict(type='sine', in_features=3, out_features=1, hidden_features=64, fusion_operator='sum'), (9, 10, 11), 'bce', 'rand'),
    ('video_st_sine_sum_h256', dict(type='sine', in_features=3, out_features=3, hidden_features=256, split_rule=[1, 2], fusion_operator='sum'), (12, 150), 'mse', 'rand'),
    # the rest of the nonlinearity table (modules.py:170-176; --model_type tanh / selu / ...)
    ('img_tanh_prod_h64', dict(type='tanh', out_features=1, hidden_features=64, fusion_operator='prod'), (14, 19), 'mse', 'linspace'),
    ('img_sigmoid_sum_h64', dict(type='sigmoid', out_features=1, hidden_features=64, fusion_operator='sum'), (11, 23), 'mse', 'rand'),
    ('img_selu_prod_h64', dict(type='selu', out_features=1, hidden_features=64, fusion_operator='prod'), (13, 16), 'mse', 'rand'),
    ('occ_elu_sum_h64', dict(type='elu', in_features=3, out_features=1, hidden_features=64, fusion_operator='sum'), (5, 6, 9), 'bce', 'rand'),
    ('img_softplus_prod_h64', dict(type='softplus', out_features=1, hidden_features=64, fusion_operator='prod'), (12, 15), 'mse', 'linspace'),
]


def run_case(name, kw, shape, loss_kind, mode, seed):
    torch.manual_seed(seed)
    kw = dict(kw)
    kw.setdefault('in_features', 2)
    kw.setdefault('mode', 'mlp')
    model = modules.SingleBVPNet(split_mlp=True, **kw)
    net = model.net
    feats_in = net.feat_per_channel if kw['mode'] != 'nerf' else \
        ([1] * kw['in_features'] if kw.get('split_rule') is None else list(kw['split_rule']))
    assert len(feats_in) == len(shape)
    coords = axes_for(shape, feats_in, seed + 1, mode)
    O = kw['out_features']
    P = int(np.prod(shape))
    g = torch.Generator().manual_seed(seed + 2)
    if loss_kind == 'mse':
        gt = torch.rand(1, P, O, generator=g) * 2 - 1
        loss_fn = lambda mo: loss_functions.image_mse(None, mo, {'img': gt})
    else:
        gt = (torch.rand(1, P, O, generator=g) < 0.4).float()
        loss_fn = lambda mo: loss_functions.occupancy_3d(mo, {'occupancy': gt})

    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    fx = {'cfg': json.dumps(dict(kw=kw, shape=list(shape), loss=loss_kind, seed=seed, name=name,
                                 feat_per_channel=list(net.feat_per_channel)))}
    for k, v in sd0.items():
        fx['sd/' + k] = v
    for c, a in enumerate(coords):
        fx[f'coords/{c}'] = a.numpy()
    fx['gt'] = gt.numpy()

    # training-style forward + backward (training.py:112-138)
    mo = model({'coords_split': coords}, ret_feat=True)
    out, hs = mo['model_out']
    fx['model_out'] = out.detach().numpy()
    for c, h in enumerate(hs):
        fx[f'feat/{c}'] = h.detach().reshape(-1, h.shape[-1]).numpy()
    losses = loss_fn({'model_out': out, 'model_in': mo['model_in']})
    total = sum(l.mean() for l in losses.values())
    fx['loss'] = np.float32(total.item())
    model.zero_grad()
    total.backward()
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            fx['grad/' + k] = prm.grad.detach().clone().numpy()
    for c, ci in enumerate(mo['model_in']):
        fx[f'dcoords/{c}'] = ci.grad.detach().clone().numpy()

    # two-stage inference API (train_img.py:125-135, train_3d_occupancy.py:128-147)
    with torch.no_grad():
        C = len(shape)
        fl = []
        for c, a in enumerate(coords):
            f = model.forward_split_channel(a.reshape(-1, a.shape[-1]), c)
            sh = [1] * C + [-1]
            sh[c] = shape[c]
            fl.append(f.reshape(sh))
        two = model.forward_split_fusion(fl)
        fx['two_stage_out'] = two.reshape(1, -1, two.shape[-1]).numpy()
        # scattered path on the dense grid (modules.py:257-309) - mlp mode with equal f_c only
        if kw['mode'] != 'nerf':
            grids = torch.meshgrid(*[torch.arange(n) for n in shape], indexing='ij')
            dense = torch.cat([coords[c].reshape(shape[c], -1)[grids[c].reshape(-1)] for c in range(C)], -1)
            fx['scattered_out'] = model({'coords': dense[None]})['model_out'].numpy()

    # 3 Adam steps (training.py:17,137-146); occupancy uses clip_grad=True
    lr = 5e-4 if loss_kind == 'bce' else 1e-4
    optim = torch.optim.Adam(lr=lr, params=model.parameters())
    loss_hist = []
    for _ in range(3):
        mo = model({'coords_split': coords})
        losses = loss_fn(mo)
        total = sum(l.mean() for l in losses.values())
        loss_hist.append(total.item())
        optim.zero_grad()
        total.backward()
        if loss_kind == 'bce':
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.)
        optim.step()
    fx['adam_lr'] = np.float32(lr)
    fx['adam_losses'] = np.array(loss_hist, np.float32)
    big = sum(v.size for v in sd0.values()) > 100_000
    for k, v in model.state_dict().items():
        if big and v.numel() > 4096:
            continue          # keep the Hd=256 fixtures small: only small tensors after Adam
        fx['adam3/' + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)
    print(f'{name}: P={P} out={tuple(out.shape)} loss={fx["loss"]:.6f}')


if __name__ == '__main__':
    import io
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])                              # optional: names of the cases to (re)generate
    for i, (name, kw, shape, loss_kind, mode) in enumerate(CASES):
        if only and name not in only:
            continue
        real, sys.stdout = sys.stdout, io.StringIO()      # the reference ctor prints the module tree
        try:
            run_case(name, kw, shape, loss_kind, mode, seed=100 + i)
            msg = sys.stdout.getvalue().strip().splitlines()[-1]
        finally:
            sys.stdout = real
        print(msg)
