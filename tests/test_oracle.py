"""Pin the numpy oracle (oracle/splitmlp_np.py) against fixtures produced by the
reference itself (oracle/gen_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest

from oracle import fixtures as FX
from oracle import splitmlp_np as O

NAMES = FX.golden_names()


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def test_fixture_inventory():
    assert len(NAMES) >= 25


@pytest.mark.parametrize('name', NAMES)
def test_forward_matches_reference(name):
    fx = FX.load(name)
    coords = [fx.encode(x) for x in fx.coords]
    out, feats = O.forward_grid(fx.params, fx.cfg, coords)
    ref = fx.data['model_out'][0]
    assert out.shape == ref.shape
    assert rel(out, ref) < 2e-5, rel(out, ref)
    for c, f in enumerate(feats):
        assert rel(f, fx.data[f'feat/{c}']) < 2e-5
    # two-stage API == forward(coords_split)  (modules.py:463-474)
    assert rel(fx.data['two_stage_out'][0], ref) < 1e-6


@pytest.mark.parametrize('name', [n for n in NAMES if 'nerf' not in n and 'fourier' not in n])
def test_scattered_path_identity(name):
    fx = FX.load(name)
    C = len(fx.shape)
    grids = np.meshgrid(*[np.arange(n) for n in fx.shape], indexing='ij')
    dense = np.concatenate([fx.coords[c][grids[c].reshape(-1)] for c in range(C)], -1)
    out = O.forward_scattered(fx.params, fx.cfg, dense)
    assert rel(out, fx.data['scattered_out'][0]) < 2e-5
    assert rel(out, fx.data['model_out'][0]) < 2e-5


@pytest.mark.parametrize('name', NAMES)
def test_loss_and_grads_match_reference(name):
    fx = FX.load(name)
    p64 = {k: v.astype(np.float64) for k, v in fx.params.items()}
    coords = [fx.encode(x.astype(np.float64)) for x in fx.coords]
    cache = {}
    out, _ = O.forward_grid(p64, fx.cfg, coords, cache)
    gt = fx.data['gt'][0].astype(np.float64)
    lval, dout = (O.image_mse if fx.loss == 'mse' else O.occupancy_3d)(out, gt)
    assert abs(lval - fx.data['loss']) < 2e-5 * max(1.0, abs(fx.data['loss']))
    grads, dcoords = O.backward_grid(p64, fx.cfg, cache, dout)
    for k, g in grads.items():
        ref = fx.data['grad/net.' + k]
        # fp32 autograd vs fp64 oracle: relative to the tensor's scale
        assert rel(g, ref) < 2e-3, (k, rel(g, ref))
    if fx.freq_bands is None:
        for c, d in enumerate(dcoords):
            ref = fx.data[f'dcoords/{c}'].reshape(d.shape)
            assert rel(d, ref) < 2e-3, (c, rel(d, ref))


@pytest.mark.parametrize('name', [n for n in NAMES if 'h256' not in n])
def test_three_adam_steps(name):
    fx = FX.load(name)
    p = {k: v.astype(np.float64).copy() for k, v in fx.params.items()}
    coords = [fx.encode(x.astype(np.float64)) for x in fx.coords]
    gt = fx.data['gt'][0].astype(np.float64)
    opt = O.Adam(p, lr=float(fx.data['adam_lr']))
    losses = []
    for _ in range(3):
        l, _, _ = O.train_step(p, fx.cfg, coords, gt, loss=fx.loss, opt=opt, clip=(fx.loss == 'bce'))
        losses.append(l)
    assert np.allclose(losses, fx.data['adam_losses'], rtol=2e-4, atol=1e-6)
    for k, v in p.items():
        ref = fx.data['adam3/net.' + k]
        # Adam's first steps move every weight by ~lr regardless of |g|: compare updates
        delta_ref = ref - fx.params[k]
        delta = v - fx.params[k].astype(np.float64)
        # (entries whose gradient is ~0 take a sign-like +-lr step that flips between fp32 and
        #  fp64 arithmetic, so a small fraction of outliers is inherent to Adam, not an error)
        bad = np.abs(delta - delta_ref) > 0.2 * float(fx.data['adam_lr']) + 1e-7
        assert bad.mean() < 0.01, (k, bad.mean())


def test_pos_enc_freq_last_is_concat_of_single_channel():
    rng = np.random.default_rng(0)
    fb = O.nerf_freq_bands(3)
    x = rng.uniform(-1, 1, size=(5, 3)).astype(np.float32)
    full = O.pos_enc(x, fb, True, True)
    parts = [O.pos_enc(x[:, c:c + 1], fb, True, True) for c in range(3)]
    assert np.array_equal(full, np.concatenate(parts, -1))
    assert full.shape[-1] == 63
    assert len(O.nerf_freq_bands(2, sidelength=512)) == 7      # 15 features @512


def test_layer_shapes_defaults():
    cfg = O.SplitCfg(in_features=2, coord_dim=2)
    assert cfg.layer_shapes() == [(256, 256)] * 3 + [(1, 256)]
    n = sum(o * i + o for o, i in cfg.layer_shapes()) + 2 * (256 + 256)
    assert n == 198657                                          # SURVEY 8a parameter count
    v = O.SplitCfg(in_features=3, coord_dim=3, out_features=3, hidden_features=1024, split_rule=[1, 2])
    n = sum(o * i + o for o, i in v.layer_shapes()) + (1024 * 1 + 1024) + (1024 * 2 + 1024)
    assert n == 3156995


@pytest.mark.parametrize('name', [n for n in NAMES if 'h256' not in n])
def test_torch_port_matches_reference(name):
    """the timing port (oracle/ref_port_torch.py) reproduces the reference's output,
    loss trajectory and 3-step Adam result."""
    import torch
    from oracle import ref_port_torch as RP
    fx = FX.load(name)
    p = RP.to_torch_params(fx.params, requires_grad=True)
    axes = [torch.as_tensor(fx.encode(x)) for x in fx.coords]
    out = RP.infer_grid(p, fx.cfg, axes, slabs=2)
    assert rel(out.reshape(-1, out.shape[-1]).numpy(), fx.data['model_out'][0]) < 1e-5
    opt = torch.optim.Adam(lr=float(fx.data['adam_lr']), params=list(p.values()))
    gt = torch.as_tensor(fx.data['gt'])
    losses = [float(RP.train_step(p, fx.cfg, axes, gt, opt, fx.loss, clip=(fx.loss == 'bce'))) for _ in range(3)]
    assert np.allclose(losses, fx.data['adam_losses'], rtol=1e-4, atol=1e-6)
