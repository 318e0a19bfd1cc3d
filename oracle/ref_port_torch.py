"""CPU ORACLE, timing flavour (test infrastructure, NOT product code).

Op-for-op torch-CPU restatement of the reference's split-MLP path, used by
`bench.py` as the `cpu_baseline` / `--impl reference` arm ("kind": "port"): the
reference IS PyTorch eager, so the same ATen calls on the same host threads are
the honest stand-in for it on a box where /root/reference does not exist.
Checked against tests/golden (tests/test_oracle.py::test_torch_port_*).

The op sequence mirrors /root/reference exactly:
  BatchLinear.forward   modules.py:26-27   matmul, then in-place += bias/num_input
  Sine.forward          modules.py:39      sin(30 * split_scale * x)   (temp + sin)
  forward_channel       modules.py:311-327
  forward_fusion        modules.py:329-354 broadcast (+|*) MATERIALISES [N_0..N_C-1, Hd]
  train step            training.py:112-146 autograd backward + torch.optim.Adam
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .splitmlp_np import SplitCfg


def to_torch_params(params, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(v).clone().float()
        out[k] = t.requires_grad_(requires_grad)
    return out


def _linear(x, W, b, num_input):
    y = x.matmul(W.permute(-1, -2))
    y += b.unsqueeze(-2) / num_input
    return y


def _act(cfg: SplitCfg, x):
    if cfg.nonlinearity == 'sine':
        return torch.sin(30 * cfg.act_scale * x)
    nl = cfg.nonlinearity
    if nl == 'relu':
        return torch.relu(x)
    fn = {'sigmoid': torch.sigmoid, 'tanh': torch.tanh, 'selu': torch.nn.functional.selu,
          'softplus': torch.nn.functional.softplus, 'elu': torch.nn.functional.elu}     # modules.py:170-176
    return fn[nl](x)


def _has_act(cfg, k):
    return k < cfg.num_hidden_layers


def forward_channel(p, cfg: SplitCfg, coord, c):
    h = _act(cfg, _linear(coord, p[f'coord_linears.{c}.weight'], p[f'coord_linears.{c}.bias'], 1))
    al = cfg.approx_layers
    for k in range(al):
        h = _linear(h, p[f'net.{k}.0.weight'], p[f'net.{k}.0.bias'], cfg.C)
        if not (k == al - 1 and (cfg.fusion_before_act or not _has_act(cfg, k))):
            h = _act(cfg, h)
    return h


def forward_fusion(p, cfg: SplitCfg, hs: Sequence[torch.Tensor]):
    h = hs[0]
    for hi in hs[1:]:
        h = h + hi if cfg.fusion_operator == 'sum' else h * hi
    if cfg.fusion_before_act:
        h = _act(cfg, h)
    if cfg.learn_code:
        h = h * p['code']
    if h.shape[-1] > cfg.fusion_feat_size:
        h = h.reshape(*h.shape[:-1], cfg.fusion_feat_size, -1).sum(-1)
    for k in range(cfg.approx_layers, cfg.num_hidden_layers + 1):
        h = _linear(h, p[f'net.{k}.0.weight'], p[f'net.{k}.0.bias'], 1)
        if _has_act(cfg, k):
            h = _act(cfg, h)
    return h


def bshape(feats: List[torch.Tensor]):
    C = len(feats)
    out = []
    for c, f in enumerate(feats):
        sh = [1] * C + [f.shape[-1]]
        sh[c] = f.shape[0]
        out.append(f.reshape(sh))
    return out


def infer_grid(p, cfg: SplitCfg, axes: Sequence[torch.Tensor], slabs: int = 1, slab_limit: int = None):
    """train_img.py:125-135 / train_3d_occupancy.py:128-147: branch features once,
    then the fusion+trunk in `slabs` slabs along axis 0 (the reference's --grid_batch).
    `slab_limit` evaluates only the first k slabs (bounded sample for timing)."""
    with torch.no_grad():
        feats = bshape([forward_channel(p, cfg, a, c) for c, a in enumerate(axes)])
        n0 = feats[0].shape[0]
        g = (n0 - 1) // slabs + 1
        outs = []
        for i in range(slabs if slab_limit is None else min(slabs, slab_limit)):
            outs.append(forward_fusion(p, cfg, [feats[0][i * g:(i + 1) * g]] + feats[1:]))
        return torch.cat(outs, 0)


def train_step(p, cfg: SplitCfg, axes, gt, optim, loss='mse', clip=False):
    """training.py:112-146 without TensorBoard; p tensors require grad."""
    feats = bshape([forward_channel(p, cfg, a, c) for c, a in enumerate(axes)])
    h = forward_fusion(p, cfg, feats)
    out = h.reshape(1, -1, h.shape[-1])
    if loss == 'mse':
        l = ((out - gt) ** 2).mean()
    else:
        l = torch.mean(torch.relu(out) - out * gt + torch.log(1 + torch.exp(-torch.abs(out))))
    optim.zero_grad()
    l.backward()
    if clip:
        torch.nn.utils.clip_grad_norm_(list(p.values()), max_norm=1.)
    optim.step()
    return l
