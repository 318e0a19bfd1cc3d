"""Host-side mirror of the reference's model API for the split-MLP path.

Same names, constructor kwargs, attributes, `state_dict` keys and forward
signatures as /root/reference/modules.py (SingleBVPNet :357-474, SplitFCBlock
:132-354, BatchLinear :12-28, Sine :31-39, PosEncodingNeRF :511-576), so
training.py / experiment_scripts/train_*.py can drive it unchanged - but the grid
path (`coords_split`, `forward_split_channel`, `forward_split_fusion`) executes in
libcoordx_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/coordx_b200.h).  PyTorch provides parameters, device memory, streams and
autograd plumbing only.

The scattered-point path `{'coords': ...}` (modules.py:257-309; SURVEY 8f "next" rank 1)
runs through the same kernels whenever no second derivative with respect to the
coordinates can be asked for (`SplitFCBlock.scattered_impl`); with grad mode on and
coordinates that require grad (summaries, diff_operators.gradient / laplace, the SDF
loss) it stays the reference's torch-op sequence.  The un-split FCBlock baseline takes the same kernels as a one-branch
split net (`_FCAsSplit`) under the same rule.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd.function import once_differentiable

from . import _lib

__all__ = ['BatchLinear', 'Sine', 'FCBlock', 'SplitFCBlock', 'SingleBVPNet', 'PosEncodingNeRF',
           'ImageDownsampling', 'sine_init', 'first_layer_sine_init', 'init_weights_normal']


# ----------------------------------------------------------------------------- layers
class BatchLinear(nn.Linear):
    """modules.py:12-28: y = x W^T, then y += bias / num_input (num_input = number of
    branches on the shared pre-fusion layers, modules.py:224-229)."""
    num_input = 1

    def forward(self, input, params=None):
        weight = self.weight if params is None else params['weight']
        bias = self.bias if params is None else params.get('bias', None)
        return F.linear(input, weight) + bias / self.num_input


class Sine(nn.Module):
    """modules.py:31-39: sin(30 * split_scale * x)."""
    split_scale = 1

    def forward(self, input):
        return torch.sin(30 * self.split_scale * input)


def init_weights_normal(m):
    if type(m) in (BatchLinear, nn.Linear) and hasattr(m, 'weight'):
        nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity='relu', mode='fan_in')


def init_weights_selu(m):
    """modules.py:619-623"""
    if type(m) in (BatchLinear, nn.Linear) and hasattr(m, 'weight'):
        nn.init.normal_(m.weight, std=1 / math.sqrt(m.weight.size(-1)))


def init_weights_elu(m):
    """modules.py:626-630"""
    if type(m) in (BatchLinear, nn.Linear) and hasattr(m, 'weight'):
        nn.init.normal_(m.weight, std=math.sqrt(1.5505188080679277) / math.sqrt(m.weight.size(-1)))


def init_weights_xavier(m):
    """modules.py:633-636"""
    if type(m) in (BatchLinear, nn.Linear) and hasattr(m, 'weight'):
        nn.init.xavier_normal_(m.weight)


def sine_init(m):
    """modules.py:639-644"""
    with torch.no_grad():
        if hasattr(m, 'weight'):
            bound = np.sqrt(6 / m.weight.size(-1)) / 30
            m.weight.uniform_(-bound, bound)


def first_layer_sine_init(m):
    """modules.py:647-652"""
    with torch.no_grad():
        if hasattr(m, 'weight'):
            bound = 1 / m.weight.size(-1)
            m.weight.uniform_(-bound, bound)


_NONLINEARITIES = {
    # name -> (factory, weight init, first-layer init)       (modules.py:170-176)
    # sine / relu run on every kernel path; the others on the FP32 kernels (`precision` resolves to fp32 for them)
    'sine': (Sine, sine_init, first_layer_sine_init),
    'relu': (lambda: nn.ReLU(inplace=False), init_weights_normal, None),
    'sigmoid': (nn.Sigmoid, init_weights_xavier, None),
    'tanh': (nn.Tanh, init_weights_xavier, None),
    'selu': (lambda: nn.SELU(inplace=False), init_weights_selu, None),
    'softplus': (nn.Softplus, init_weights_normal, None),
    'elu': (lambda: nn.ELU(inplace=False), init_weights_elu, None),
}


def _layer(in_f, out_f, nl=None):
    mods = [BatchLinear(in_f, out_f)]
    if nl is not None:
        mods.append(nl)
    return nn.Sequential(*mods)


# ----------------------------------------------------------------------------- C-ABI glue
def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _as_f32c(t):
    return t.detach().to(torch.float32).contiguous()


class _NetDesc:
    """Builds `struct coordx_net` for a SplitFCBlock; keeps the tensors it points to alive."""

    def __init__(self, block: 'SplitFCBlock', params, n_tables=None):
        self.keep = []
        n = _lib.Net()
        # the trunk only needs the number of feature tables being fused, which may differ from the
        # branch count when the caller pre-combined some axes (train_video.py:121-128)
        n.n_branches = block.split_channels if n_tables is None else n_tables
        for c, f in enumerate(block.feat_per_channel):
            n.branch_in[c] = f
        layers = block.layer_dims
        n.pre_width = block.pre_width
        n.fuse_width = block.fuse_width
        n.hidden = block.fusion_feat_size
        n.n_pre = block.approx_layers
        n.n_layers = len(layers)
        for k, (o, i) in enumerate(layers):
            n.layer_in[k], n.layer_out[k] = i, o
        n.act = _lib.ACTS[block.nonlinearity]
        n.fusion_op = {'sum': _lib.FUSE_SUM, 'prod': _lib.FUSE_PROD}[block.fusion_operator]
        n.fusion_before_act = int(block.fusion_before_act)
        n.has_code = int(block.learn_code)
        n.act_scale = float(block.act_scale)
        it = iter(params)
        for c in range(block.split_channels):
            n.first_w[c] = self._p(next(it))
            n.first_b[c] = self._p(next(it))
        for k in range(len(layers)):
            n.layer_w[k] = self._p(next(it))
            n.layer_b[k] = self._p(next(it))
        if block.learn_code:
            n.code = self._p(next(it))
        self.net = n

    def _p(self, t):
        t = _as_f32c(t)
        if not t.is_cuda:
            raise _lib.CoordxError('coordx_b200: parameters must live on a CUDA device (model.cuda()); '
                                   'there is no CPU path for the split-MLP grid kernels')
        self.keep.append(t)
        return t.data_ptr()


def _zeros_like_flat(tensors):
    """fp32 zero tensors shaped like `tensors`, carved out of ONE zero-filled allocation (a single fill kernel;
    every piece starts 16-byte aligned)."""
    sizes = [(int(t.numel()) + 3) // 4 * 4 for t in tensors]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=tensors[0].device)
    out, off = [], 0
    for t, n in zip(tensors, sizes):
        out.append(flat[off:off + t.numel()].view(t.shape))
        off += n
    return out


def _grads_for(block, params):
    g = _lib.Grads()
    bufs = _zeros_like_flat(params)            # one fill for all of them, not one per parameter
    it = iter(bufs)
    for c in range(block.split_channels):
        g.first_w[c] = next(it).data_ptr()
        g.first_b[c] = next(it).data_ptr()
    for k in range(len(block.layer_dims)):
        g.layer_w[k] = next(it).data_ptr()
        g.layer_b[k] = next(it).data_ptr()
    if block.learn_code:
        g.code = next(it).data_ptr()
    return g, bufs


_WS = {}
_WS_KEEP_BYTES = int(float(os.environ.get('COORDX_WS_KEEP_MB', '1024')) * (1 << 20))


def _workspace(nbytes, device):
    """Scratch buffer for one C call.  Buffers up to COORDX_WS_KEEP_MB (default 1 GiB) are cached per (device, stream)
    and grow on demand: kernels that use them are ordered on that stream, so consecutive calls may share one (the C
    library never keeps pointers into it across calls).  Larger requests (the stored dZ tiles of a big training
    grid: 2 GB for 2048^2, 8 GB for 4096^2) are plain allocations that go back to torch's caching allocator when the
    call returns, so `torch.cuda.empty_cache()` can hand the memory back to the driver."""
    nbytes = max(int(nbytes), 256)
    if nbytes > _WS_KEEP_BYTES or torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def release_workspaces():
    """Drop every cached scratch buffer (they are re-created on demand)."""
    _WS.clear()


def _backward_workspace(lib, net, n_arr, n0, prec, device):
    """Workspace of coordx_trunk_backward: the preferred size (the tcgen05 path keeps its dZ tiles between its two
    passes), or - when that allocation fails - the minimal one, with which the library recomputes instead."""
    want = lib.coordx_trunk_workspace_bytes(net, n_arr, 0, n0, prec, 1)
    try:
        return _workspace(want, device)
    except torch.cuda.OutOfMemoryError:
        small = lib.coordx_trunk_workspace_bytes(net, n_arr, 0, n0, prec, 2)
        if small >= want:
            raise
        release_workspaces()
        torch.cuda.empty_cache()
        return _workspace(small, device)


def _net_desc(block, params, n_tables=None):
    """struct coordx_net for `block`, rebuilt only when a parameter tensor moved."""
    if not all(p.dtype == torch.float32 and p.is_contiguous() for p in params):
        return _NetDesc(block, params, n_tables)          # _as_f32c makes copies: nothing to cache
    key = tuple(p.data_ptr() for p in params)
    cache = block.__dict__.setdefault('_desc_cache', {})
    d = cache.get(n_tables)
    if d is None or d.key != key:
        d = _NetDesc(block, params, n_tables)
        d.key = key
        cache[n_tables] = d
    return d


def _branch_forward(block, branch, coords, params, save=False):
    """forward_channel through the C ABI.  save=True (training): the branch activations are kept in a buffer of
    their own for coordx_branch_backward_saved instead of being recomputed by the backward pass."""
    lib = _lib.load()
    x = _as_f32c(coords)
    if not x.is_cuda:
        raise _lib.CoordxError('coordx_b200: the grid path needs CUDA tensors (no CPU fallback)')
    n = x.shape[0]
    desc = _net_desc(block, params)
    feat = torch.empty(n, block.fuse_width, dtype=torch.float32, device=x.device)
    saved = None
    with torch.cuda.device(x.device):
        if save:
            ws_bytes = lib.coordx_branch_workspace_bytes(desc.net, branch, n, 1)
            saved = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=x.device)
            _lib.check(lib.coordx_branch_forward_save(desc.net, branch, x.data_ptr(), n, feat.data_ptr(), saved.data_ptr(),
                                                      saved.numel(), _stream_ptr()), 'coordx_branch_forward_save')
        else:
            ws_bytes = lib.coordx_branch_workspace_bytes(desc.net, branch, n, 0)
            ws = _workspace(ws_bytes, x.device)
            _lib.check(lib.coordx_branch_forward(desc.net, branch, x.data_ptr(), n, feat.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), _stream_ptr()), 'coordx_branch_forward')
    return feat, x, saved


def _fusion_forward(block, precision, feats, params):
    lib = _lib.load()
    feats = [_as_f32c(t) for t in feats]
    dev = feats[0].device
    if not dev.type == 'cuda':
        raise _lib.CoordxError('coordx_b200: the grid path needs CUDA tensors (no CPU fallback)')
    ns = [f.shape[0] for f in feats]
    desc = _net_desc(block, params, len(feats))
    prec = block.resolve_precision(desc.net, precision)
    P = 1
    for v in ns:
        P *= v
    out = torch.empty(P, block.out_width, dtype=torch.float32, device=dev)
    n_arr = _lib.i64_array(ns)
    f_arr = _lib.ptr_array([f.data_ptr() for f in feats])
    # SingleBVPNet.forward_with_loss parks (kind, gt) here: the trunk call then also evaluates the loss of the training
    # step (coordx_trunk_forward_loss; in the tcgen05 kernel's epilogue) and leaves its value and d loss / d out behind
    hook = getattr(block, '_loss_hook', None)
    if hook is not None and (hook.get('loss') is not None or hook['gt'].numel() != out.numel()):
        hook = None                                  # already served by an earlier trunk call / not this grid
    with torch.cuda.device(dev):
        ws_bytes = lib.coordx_trunk_workspace_bytes(desc.net, n_arr, 0, ns[0], prec, 0)
        ws = _workspace(ws_bytes, dev)
        if hook is None:
            _lib.check(lib.coordx_trunk_forward(desc.net, f_arr, n_arr, 0, ns[0], out.data_ptr(), prec, ws.data_ptr(),
                                                ws.numel(), _stream_ptr()), 'coordx_trunk_forward')
        else:
            gt = _as_f32c(hook['gt'].reshape(P, block.out_width))
            loss = torch.zeros((), dtype=torch.float32, device=dev)
            dout = torch.empty_like(out)
            _lib.check(lib.coordx_trunk_forward_loss(desc.net, f_arr, n_arr, 0, ns[0], out.data_ptr(), hook['kind'],
                                                     gt.data_ptr(), 1.0 / max(out.numel(), 1), loss.data_ptr(),
                                                     dout.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream_ptr()),
                       'coordx_trunk_forward_loss')
            hook['loss'], hook['dout'] = loss, dout
    return out, feats, prec


class _TrunkLossFn(torch.autograd.Function):
    """The loss value the trunk call has already produced (coordx_trunk_forward_loss), attached to `out` in the
    autograd graph: its backward hands the saved d loss / d out to the trunk's backward.  unit_grad: the caller
    promises an upstream gradient of exactly 1 (`loss.backward()` on this scalar), see loss_functions.fused_loss."""

    @staticmethod
    def forward(ctx, out, loss, dout, unit_grad):
        ctx.dout, ctx.shape, ctx.unit_grad = dout, out.shape, unit_grad
        return loss.detach()

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        d = ctx.dout.reshape(ctx.shape)
        return (d if ctx.unit_grad else d * g), None, None, None


def _needs_grad(tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


class _BranchFn(torch.autograd.Function):
    """SplitFCBlock.forward_channel (modules.py:311-327) through coordx_branch_forward/backward."""

    @staticmethod
    def forward(ctx, block, branch, coords, *params):
        feat, x, saved = _branch_forward(block, branch, coords, params, save=True)
        ctx.block, ctx.branch, ctx.saved_acts = block, branch, saved
        ctx.save_for_backward(x, *params)
        return feat

    @staticmethod
    @once_differentiable
    def backward(ctx, dfeat):
        lib = _lib.load()
        x, *params = ctx.saved_tensors
        block, branch = ctx.block, ctx.branch
        n = x.shape[0]
        desc = _NetDesc(block, params)
        grads, bufs = _grads_for(block, params)
        d = _as_f32c(dfeat)
        dcoords = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(x.device):
            ws = ctx.saved_acts                     # written by coordx_branch_forward_save
            _lib.check(lib.coordx_branch_backward_saved(desc.net, branch, x.data_ptr(), n, d.data_ptr(), grads,
                                                        dcoords.data_ptr() if dcoords is not None else None,
                                                        ws.data_ptr(), ws.numel(), _stream_ptr()),
                       'coordx_branch_backward_saved')
        used = block.branch_param_mask(branch)
        out = [b if (u and ctx.needs_input_grad[3 + i]) else None for i, (b, u) in enumerate(zip(bufs, used))]
        return (None, None, dcoords, *out)


class _FusionFn(torch.autograd.Function):
    """SplitFCBlock.forward_fusion (modules.py:329-354) through coordx_trunk_forward/backward.
    feats: C tables [N_c, F]; returns [prod N_c, out_width]."""

    @staticmethod
    def forward(ctx, block, precision, nfeat, *tensors):
        params = tensors[nfeat:]
        out, feats, prec = _fusion_forward(block, precision, tensors[:nfeat], params)
        ctx.block, ctx.prec, ctx.nfeat = block, prec, nfeat
        ctx.save_for_backward(*feats, *params)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        lib = _lib.load()
        block, nfeat = ctx.block, ctx.nfeat
        feats = list(ctx.saved_tensors[:nfeat])
        params = ctx.saved_tensors[nfeat:]
        dev = feats[0].device
        ns = [f.shape[0] for f in feats]
        desc = _NetDesc(block, params, n_tables=nfeat)
        grads, bufs = _grads_for(block, params)
        d = _as_f32c(dout)
        dfeats = _zeros_like_flat(feats)
        n_arr = _lib.i64_array(ns)
        f_arr = _lib.ptr_array([f.data_ptr() for f in feats])
        df_arr = _lib.ptr_array([f.data_ptr() for f in dfeats])
        with torch.cuda.device(dev):
            ws = _backward_workspace(lib, desc.net, n_arr, ns[0], ctx.prec, dev)
            _lib.check(lib.coordx_trunk_backward(desc.net, f_arr, n_arr, 0, ns[0], d.data_ptr(), grads, df_arr, ctx.prec,
                                                 ws.data_ptr(), ws.numel(), _stream_ptr()), 'coordx_trunk_backward')
        used = block.trunk_param_mask()
        pg = [b if (u and ctx.needs_input_grad[3 + nfeat + i]) else None for i, (b, u) in enumerate(zip(bufs, used))]
        fg = [df if ctx.needs_input_grad[3 + i] else None for i, df in enumerate(dfeats)]
        return (None, None, None, *fg, *pg)


# ----------------------------------------------------------------------------- blocks
class _FCAsSplit:
    """An un-split FCBlock seen through the C ABI: a split net with ONE branch (`coord_linears[0]` = the first layer,
    every hidden layer but the last = the branch's shared layers) whose "fusion" multiplies by a one-row table of
    ones, followed by the trunk (last hidden layer + output layer).  Same arithmetic as `FCBlock.net(coords)`
    (BatchLinear + activation per layer, modules.py:42-107); it lets the baseline the paper's speed-up is defined
    against (train_img.py:115-123) run on the same kernels: per-layer 3xTF32 GEMMs over the P points, then the
    tcgen05 trunk."""

    def __init__(self, fc: 'FCBlock'):
        lins = [seq[0] for seq in fc.net]
        self.split_channels = 1
        self.feat_per_channel = [lins[0].in_features]
        self.first, self.layers = lins[0], lins[1:]
        self.layer_dims = [(l.out_features, l.in_features) for l in self.layers]
        self.approx_layers = max(len(self.layers) - 2, 0)           # all hidden layers but one are "pre-fusion"
        self.pre_width = lins[0].out_features
        self.fuse_width = self.layer_dims[self.approx_layers - 1][0] if self.approx_layers > 0 else self.pre_width
        self.fusion_feat_size = self.fuse_width
        self.out_width = self.layer_dims[-1][0]
        self.nonlinearity = fc.nonlinearity
        self.fusion_operator, self.fusion_before_act, self.learn_code, self.act_scale = 'prod', False, False, 1.0
        self.precision = fc.precision

    def abi_params(self):
        ps = [self.first.weight, self.first.bias]
        for l in self.layers:
            ps += [l.weight, l.bias]
        return ps

    def branch_param_mask(self, branch):
        return [True, True] + [k < self.approx_layers for k in range(len(self.layers)) for _ in (0, 1)]

    def trunk_param_mask(self):
        return [False, False] + [k >= self.approx_layers for k in range(len(self.layers)) for _ in (0, 1)]

    def resolve_precision(self, net, precision=None):
        return SplitFCBlock.resolve_precision(self, net, precision)


class FCBlock(nn.Module):
    """Un-split MLP baseline (modules.py:42-129).  `fused_impl`: 'auto' (default) runs CUDA inputs through the split
    path's kernels (see _FCAsSplit) unless gradients with respect to the coordinates may be asked for; 'eager' is the
    reference's nn.Sequential."""

    def __init__(self, in_features, out_features, num_hidden_layers, hidden_features, outermost_linear=False,
                 nonlinearity='relu', weight_init=None, approx_layers=2, fusion_size=1, reduced=False):
        super().__init__()
        factory, nl_init, first_init = _NONLINEARITIES[nonlinearity]
        nl = factory()
        self.weight_init = weight_init if weight_init is not None else nl_init
        s = 1 if reduced else fusion_size
        in_size = out_size = hidden_features * s
        layers = [_layer(in_features, out_size, nl)]
        for i in range(num_hidden_layers):
            if i + 1 == approx_layers:
                out_size = hidden_features if not reduced else hidden_features * fusion_size
            if i == approx_layers:
                out_size = hidden_features
            layers.append(_layer(in_size, out_size, nl))
            in_size = out_size
        layers.append(_layer(in_size, out_features, None if outermost_linear else nl))
        self.net = nn.Sequential(*layers)
        self.net.apply(self.weight_init)
        if first_init is not None:
            self.net[0].apply(first_init)
        self.nonlinearity, self.precision, self.fused_impl = nonlinearity, None, 'auto'
        # representable as a one-branch split net: sine / relu, linear output layer, at least one hidden layer after
        # the first, no group-sum widths
        self._fusable = (nonlinearity in ('sine', 'relu') and outermost_linear and num_hidden_layers >= 1 and
                         fusion_size == 1 and not reduced)

    def forward(self, coords, params=None, **kwargs):
        if self.fused_impl not in ('auto', 'eager'):
            raise ValueError(f'unknown fused_impl {self.fused_impl!r}')
        if (self.fused_impl == 'auto' and self._fusable and coords.is_cuda and coords.numel() > 0 and
                not (torch.is_grad_enabled() and coords.requires_grad)):
            v = self.__dict__.get('_split_view')
            if v is None:
                v = self.__dict__['_split_view'] = _FCAsSplit(self)
            v.precision = self.precision
            lead = coords.shape[:-1]
            x = coords.reshape(-1, coords.shape[-1])
            ps = v.abi_params()
            h = _BranchFn.apply(v, 0, x, *ps) if _needs_grad([x, *ps]) else _branch_forward(v, 0, x, ps)[0]
            neutral = torch.ones(1, h.shape[-1], dtype=h.dtype, device=h.device)
            if _needs_grad([h, *ps]):
                out = _FusionFn.apply(v, None, 2, neutral, h, *ps)
            else:
                out = _fusion_forward(v, None, [neutral, h], ps)[0]
            return out.reshape(*lead, out.shape[-1])
        return self.net(coords)


class SplitFCBlock(nn.Module):
    """modules.py:132-354.  Same constructor / attributes / parameter names; the grid path
    runs in libcoordx_b200.so."""

    def __init__(self, in_features, out_features, num_hidden_layers, hidden_features, outermost_linear=False,
                 nonlinearity='relu', weight_init=None, coord_dim=2, approx_layers=2, split_rule=None,
                 fusion_operator='sum', act_scale=1, fusion_before_act=False, use_atten=False, learn_code=False,
                 last_layer_features=-1, fusion_size=1, reduced=False, precision='auto'):
        super().__init__()
        if nonlinearity not in _NONLINEARITIES:
            raise KeyError(nonlinearity)                       # modules.py:179
        if fusion_operator not in ('sum', 'prod'):
            # the reference silently skips the fusion for unknown strings (modules.py:336-341)
            raise ValueError(f"fusion_operator must be 'sum' or 'prod', got {fusion_operator!r}")
        if not outermost_linear:
            raise NotImplementedError('coordx_b200 covers outermost_linear=True (what SingleBVPNet builds, modules.py:406)')
        if use_atten:
            raise NotImplementedError('use_atten is ignored by the reference grid path and is out of scope here')
        self.first_layer_init = None
        self.fork_branches = os.environ.get('COORDX_FORK_BRANCHES', '1') != '0'
        self._branch_streams = []
        self.coord_dim = coord_dim
        per = in_features // coord_dim
        self.feat_per_channel = [per] * coord_dim if split_rule is None else [per * k for k in split_rule]
        self.split_channels = len(self.feat_per_channel)
        if self.split_channels > _lib.MAX_BRANCHES:
            raise ValueError('too many branches')
        self.approx_layers = approx_layers
        self.num_hidden_layers = num_hidden_layers
        self.module_prefix = ""
        self.fusion_operator = fusion_operator
        self.fusion_before_act = fusion_before_act
        self.out_features = out_features
        self.use_atten = use_atten
        self.learn_code = learn_code
        self.nonlinearity = nonlinearity
        self.act_scale = act_scale
        self.precision = precision
        self.fusion_size = 1
        self.fusion_feat_size = out_features
        if approx_layers != num_hidden_layers + 1:
            last_layer_features = 1
            self.fusion_size = fusion_size
            self.fusion_feat_size = hidden_features
        elif last_layer_features < 0:
            last_layer_features = hidden_features
        last_layer_features = last_layer_features * out_features

        factory, nl_init, first_init = _NONLINEARITIES[nonlinearity]
        nl = factory()
        self.weight_init = weight_init if weight_init is not None else nl_init

        s = 1 if reduced else fusion_size
        self.pre_width = hidden_features * s
        self.coord_linears = nn.ModuleList([BatchLinear(f, self.pre_width) for f in self.feat_per_channel])
        self.coord_nl = nl
        self.coord_nl.split_scale = act_scale
        self.coord_linears.apply(first_init if first_init is not None else self.weight_init)

        dims = []                                              # (out, in) of net[k]
        n_shared = max(min(approx_layers, num_hidden_layers) - 1, 0)
        dims += [(self.pre_width, self.pre_width)] * n_shared
        dims.append((hidden_features * fusion_size, self.pre_width))
        dims += [(hidden_features, hidden_features)] * (num_hidden_layers - n_shared - 1)
        dims.append((last_layer_features, hidden_features))
        self.layer_dims = dims
        self.net = nn.Sequential(*[_layer(i, o, nl if k < len(dims) - 1 else None) for k, (o, i) in enumerate(dims)])
        if self.weight_init is not None:
            self.net.apply(self.weight_init)
        for k in range(min(approx_layers, len(dims))):
            self.net[k][0].num_input = self.split_channels
        if learn_code:
            self.code = nn.Parameter(torch.ones(hidden_features * fusion_size))
        self.fuse_width = dims[approx_layers - 1][0] if approx_layers > 0 else self.pre_width
        # approx_layers == num_hidden_layers + 1: no trunk, the group-sum of modules.py:349-351 is the output
        self.out_width = last_layer_features if approx_layers < len(dims) else self.fusion_feat_size
        if len(dims) > _lib.MAX_LAYERS:
            raise ValueError('too many layers')

    # ---- parameter plumbing ------------------------------------------------------
    def abi_params(self):
        ps = []
        for lin in self.coord_linears:
            ps += [lin.weight, lin.bias]
        for seq in self.net:
            ps += [seq[0].weight, seq[0].bias]
        if self.learn_code:
            ps.append(self.code)
        return ps

    def branch_param_mask(self, branch):
        m = []
        for c in range(self.split_channels):
            m += [c == branch] * 2
        for k in range(len(self.layer_dims)):
            m += [k < self.approx_layers] * 2
        if self.learn_code:
            m.append(False)
        return m

    def trunk_param_mask(self):
        m = [False] * (2 * self.split_channels)
        for k in range(len(self.layer_dims)):
            m += [k >= self.approx_layers] * 2
        if self.learn_code:
            m.append(True)
        return m

    def resolve_precision(self, net, precision=None):
        p = precision or self.precision or 'auto'
        if p == 'auto':
            p = os.environ.get('COORDX_PRECISION', 'auto')
        lib = _lib.load()
        if p == 'auto':
            # SIREN activations live in [-1, 1]: bf16 operands meet the 2e-3 output bar; ReLU nets are
            # unnormalised, fp16's 3 extra mantissa bits are needed there (tests/test_gpu_parity.py)
            want = _lib.PREC_BF16 if self.nonlinearity == 'sine' else _lib.PREC_FP16
            return want if lib.coordx_trunk_tc_supported(net, want) else _lib.PREC_FP32
        code = _lib.PRECISIONS[p]
        if code != _lib.PREC_FP32 and not lib.coordx_trunk_tc_supported(net, code):
            raise _lib.CoordxError(f'precision={p!r}: the tcgen05 trunk kernel does not cover this architecture '
                                   '(needs hidden_features=256, one hidden trunk layer, fusion_size=1); use fp32/auto')
        return code

    # ---- the hot path ------------------------------------------------------------
    def forward_channel(self, coord, channel_id, pos_codes=None):
        """modules.py:311-327.  coord [..., f_c] -> [..., F]"""
        lead = coord.shape[:-1]
        params = self.abi_params()
        x2 = coord.reshape(-1, coord.shape[-1])
        if _needs_grad([coord, *params]):
            h = _BranchFn.apply(self, channel_id, x2, *params)
        else:
            h = _branch_forward(self, channel_id, x2, params)[0]        # inference: no autograd bookkeeping
        h = h.reshape(*lead, h.shape[-1])
        if pos_codes is not None and self.approx_layers > 0:
            h = h + pos_codes
        return h

    def forward_fusion(self, hs, precision=None):
        """modules.py:329-354.  hs: iterable of broadcastable feature tensors [..., F]."""
        hs = list(hs)
        plan = _broadcast_axes([tuple(h.shape[:-1]) for h in hs])
        if plan is None:
            leads = {tuple(h.shape[:-1]) for h in hs}
            if len(leads) == 1 and len(hs) >= 1 and hs[0].is_cuda and hs[0].numel() > 0:
                # Same leading shape everywhere = per-point features, no broadcast: the interpolated branch features
                # of the rendering path (ray_rendering.py:139-144 hands forward_split_fusion a stacked [C, .., Hd]
                # tensor).  One elementwise pass fuses them; the trunk then sees a 1 x P grid whose first axis is
                # the fusion's neutral element, i.e. the same kernels as the grid (see _forward_scattered_fused).
                h = hs[0]
                for hi in hs[1:]:
                    h = h + hi if self.fusion_operator == 'sum' else h * hi
                lead = h.shape[:-1]
                h2 = h.reshape(-1, h.shape[-1])
                neutral = (torch.zeros if self.fusion_operator == 'sum' else torch.ones)(1, h2.shape[-1], dtype=h2.dtype,
                                                                                        device=h2.device)
                params = self.abi_params()
                if _needs_grad([h2, *params]):
                    out = _FusionFn.apply(self, precision, 2, neutral, h2, *params)
                else:
                    out = _fusion_forward(self, precision, [neutral, h2], params)[0]
                return out.reshape(*lead, out.shape[-1])
            return self._fusion_eager(hs)
        order, ns, out_lead = plan
        feats = [hs[i].reshape(-1, hs[i].shape[-1]) for i in order]
        if len(feats) < 2 or len(feats) > _lib.MAX_BRANCHES:
            return self._fusion_eager(hs)
        params = self.abi_params()
        if _needs_grad([*feats, *params]):
            out = _FusionFn.apply(self, precision, len(feats), *feats, *params)
        else:
            out = _fusion_forward(self, precision, feats, params)[0]
        return out.reshape(*out_lead, out.shape[-1])

    def _fusion_eager(self, hs):
        """Broadcast patterns that are not an outer product of per-tensor axes (e.g. a shared
        batch dimension > 1): plain torch ops, as the reference."""
        h = hs[0]
        for hi in hs[1:]:
            h = h + hi if self.fusion_operator == 'sum' else h * hi
        if self.fusion_before_act:
            h = self.net[self.approx_layers - 1][1](h)
        if self.learn_code:
            h = h * self.code
        if h.shape[-1] > self.fusion_feat_size:
            h = h.reshape(*h.shape[:-1], self.fusion_feat_size, -1).sum(-1)
        for k in range(self.approx_layers, self.num_hidden_layers + 1):
            h = self.net[k](h)
        return h

    def forward(self, coords, params=None, pos_codes=None, split_coord=False, ret_feat=False, **kwargs):
        if split_coord:                                        # modules.py:248-255
            if self.fork_branches and len(coords) > 1 and coords[0].is_cuda:
                # the per-axis branch MLPs are independent: one stream per branch, joined before the fusion.  Autograd
                # runs each backward on its forward's stream, so the branch backwards overlap too (and a captured
                # training step gets parallel graph branches)
                main = torch.cuda.current_stream(coords[0].device)
                if len(self._branch_streams) < len(coords):
                    self._branch_streams = [torch.cuda.Stream(coords[0].device) for _ in coords]
                    # parameter gradients produced on a branch stream are accumulated on the main stream: intended
                    # (autograd inserts the synchronisation), not worth a warning per backward
                    warn_off = getattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch', None)
                    if warn_off is not None:
                        warn_off(False)
                hs = []
                for i, coord in enumerate(coords):
                    st = self._branch_streams[i]
                    st.wait_stream(main)
                    with torch.cuda.stream(st):
                        h_i = self.forward_channel(coord, i, pos_codes)
                    h_i.record_stream(main)
                    hs.append(h_i)
                for st in self._branch_streams[:len(coords)]:
                    main.wait_stream(st)
            else:
                hs = [self.forward_channel(coord, i, pos_codes) for i, coord in enumerate(coords)]
            h = self.forward_fusion(hs)
            out = h.reshape(h.shape[0], -1, h.shape[-1])
            return (out, hs) if ret_feat else out
        return self._forward_scattered(coords, pos_codes, ret_feat)

    def _forward_scattered_fused(self, coords, ret_feat=False):
        """modules.py:257-309 through the same kernels as the grid path: every query point is one 'axis sample' of
        every branch (C branch-MLP calls over P rows: the per-layer 3xTF32 GEMM chain for large P), the per-point
        fusion is one elementwise pass, and the trunk sees a 1 x P grid whose first axis is the fusion's neutral
        element - so the tcgen05 trunk (or the separable sum evaluation) serves scattered points unchanged."""
        lead = coords.shape[:-1]
        x = coords.reshape(-1, coords.shape[-1])
        parts = torch.split(x, self.feat_per_channel, dim=-1)
        feats = [self.forward_channel(p_.contiguous(), c) for c, p_ in enumerate(parts)]       # C x [P, F]
        h = feats[0]
        for f_ in feats[1:]:
            h = h + f_ if self.fusion_operator == 'sum' else h * f_
        neutral = (torch.zeros if self.fusion_operator == 'sum' else torch.ones)(1, h.shape[-1], dtype=h.dtype, device=h.device)
        params = self.abi_params()
        if _needs_grad([h, *params]):
            out = _FusionFn.apply(self, None, 2, neutral, h, *params)
        else:
            out = _fusion_forward(self, None, [neutral, h], params)[0]
        out = out.reshape(*lead, out.shape[-1])
        if ret_feat:
            return out, torch.stack([f_.reshape(*lead, f_.shape[-1]) for f_ in feats], -2)
        return out

    def _forward_scattered(self, coords, pos_codes=None, ret_feat=False):
        """Arbitrary points [..., sum f_c] (modules.py:257-309).  `scattered_impl`: 'auto' (default) takes the CUDA
        kernels unless gradients with respect to the coordinates may be asked for (grad mode on and
        coords.requires_grad: diff_operators.gradient / laplace differentiate twice, which only the torch-op
        formulation supports); 'fused' always takes the kernels (first derivatives only; a second differentiation
        raises); 'eager' is the reference's op sequence."""
        impl = getattr(self, 'scattered_impl', 'auto')
        if impl not in ('auto', 'fused', 'eager'):
            raise ValueError(f'unknown scattered_impl {impl!r}')
        wants_coord_grad = torch.is_grad_enabled() and coords.requires_grad
        if coords.is_cuda and pos_codes is None and coords.numel() > 0 and 2 <= _lib.MAX_BRANCHES and \
                (impl == 'fused' or (impl == 'auto' and not wants_coord_grad)):
            return self._forward_scattered_fused(coords, ret_feat)
        al = self.approx_layers
        parts = torch.split(coords, self.feat_per_channel, dim=-1)
        hs = torch.stack([lin(x) for lin, x in zip(self.coord_linears, parts)], -2)
        hs = self.coord_nl(hs)
        for k in range(al - 1):
            hs = self.net[k](hs)
        if al > 0:
            hs = self.net[al - 1][0](hs) if self.fusion_before_act else self.net[al - 1](hs)
        if pos_codes is not None:
            hs = hs + pos_codes
        h = hs.sum(-2) if self.fusion_operator == 'sum' else hs.prod(-2)
        if self.fusion_before_act:
            h = self.net[al - 1][1](h)
        if self.learn_code:
            h = h * self.code
        h = h.reshape(*h.shape[:-1], self.fusion_feat_size, -1).sum(-1)
        for k in range(al, self.num_hidden_layers + 1):
            h = self.net[k](h)
        return (h, hs) if ret_feat else h


def _broadcast_axes(shapes):
    """Decide whether broadcasting the given leading shapes is an outer product in which every
    tensor owns a disjoint, contiguous run of output dims.  Returns (tensor order, axis lengths,
    output leading shape) or None."""
    nd = max((len(s) for s in shapes), default=0)
    padded = [(1,) * (nd - len(s)) + tuple(s) for s in shapes]
    out_lead = [1] * nd
    owner = [None] * nd
    spans = []
    for t, s in enumerate(padded):
        dims = [d for d in range(nd) if s[d] != 1]
        for d in dims:
            if owner[d] is not None:
                return None
            owner[d] = t
            out_lead[d] = s[d]
        if dims and dims != list(range(dims[0], dims[-1] + 1)):
            return None
        spans.append((dims[0] if dims else -1, t, int(np.prod([s[d] for d in dims])) if dims else 1))
    spans.sort()
    order = [t for _, t, _ in spans]
    ns = [n for _, _, n in spans]
    return order, ns, out_lead


# ----------------------------------------------------------------------------- model
class ImageDownsampling(nn.Module):
    """modules.py:477-508; never enabled by any script (downsample=False everywhere)."""

    def __init__(self, sidelength, downsample=False):
        super().__init__()
        self.sidelength = (sidelength, sidelength) if isinstance(sidelength, int) else sidelength
        if self.sidelength is None:
            assert downsample is False
        if downsample:
            raise NotImplementedError('downsample=True is out of scope')
        self.downsample = downsample

    def forward(self, coords):
        return coords


class PosEncodingNeRF(nn.Module):
    """modules.py:511-576: [x, sin(f_k x), cos(f_k x)]_k, per-axis contiguous when freq_last."""

    def __init__(self, in_features, sidelength=None, fn_samples=None, use_nyquist=True, freq_last=False,
                 freq_params=None, include_coord=True):
        super().__init__()
        self.in_features = in_features
        self.freq_last = freq_last
        self.include_coord = include_coord
        if freq_params is None:
            if in_features == 3:
                self.num_frequencies = 10
            elif in_features == 2:
                assert sidelength is not None
                if isinstance(sidelength, int):
                    sidelength = (sidelength, sidelength)
                self.num_frequencies = self.get_num_frequencies_nyquist(min(sidelength)) if use_nyquist else 4
            elif in_features == 1:
                if fn_samples is None:
                    fn_samples = min(sidelength[0], sidelength[1])
                self.num_frequencies = self.get_num_frequencies_nyquist(fn_samples) if use_nyquist else 4
            else:
                raise ValueError(in_features)
            bands = 2 ** torch.arange(self.num_frequencies) * np.pi
        else:
            self.num_frequencies = int(freq_params[1])
            bands = 2 * (2. ** torch.linspace(0, freq_params[0], self.num_frequencies) - 1) * np.pi
        self.out_dim = in_features * (2 * self.num_frequencies + (1 if include_coord else 0))
        self.freq_bands = nn.Parameter(bands.float(), requires_grad=False)

    @staticmethod
    def get_num_frequencies_nyquist(samples):
        return int(math.floor(math.log(1 / (2 * (2 * 1 / samples)), 2)))

    def forward(self, coords, single_channel=False):
        f = coords.shape[-1] if single_channel else self.in_features
        ang = coords.unsqueeze(-2) * self.freq_bands.reshape([1] * (coords.dim() - 1) + [-1, 1])
        enc = torch.cat([torch.sin(ang), torch.cos(ang)], -1).flatten(-2)
        if self.include_coord:
            enc = torch.cat([coords, enc], -1)
        if self.freq_last:
            lead = enc.shape[:-1]
            enc = enc.reshape(*lead, -1, f).transpose(-1, -2).reshape(*lead, -1)
        return enc


class SingleBVPNet(nn.Module):
    """modules.py:357-474.  `precision` ('auto' | 'fp32' | 'bf16' | 'fp16') and `quiet` are the
    only additions to the reference kwargs."""

    def __init__(self, out_features=1, type='sine', in_features=2, mode='mlp', hidden_features=256,
                 num_hidden_layers=3, split_mlp=False, split_rule=None, pos_enc=False, **kwargs):
        super().__init__()
        quiet = kwargs.pop('quiet', False)
        precision = kwargs.pop('precision', 'auto')
        if not quiet:
            print(kwargs)
        self.mode = mode
        self.split_mlp = split_mlp
        self.pos_enc = pos_enc
        self.module_prefix = ""
        coord_dim = in_features
        if mode == 'nerf':
            self.positional_encoding = PosEncodingNeRF(in_features=in_features, sidelength=kwargs.get('sidelength', None),
                                                       fn_samples=kwargs.get('fn_samples', None),
                                                       use_nyquist=kwargs.get('use_nyquist', True),
                                                       freq_params=kwargs.get('freq_params', None),
                                                       include_coord=kwargs.get('include_coord', True),
                                                       freq_last=split_mlp)
            in_features = self.positional_encoding.out_dim
        elif mode != 'mlp':
            raise NotImplementedError(f"mode={mode!r} is outside the split-MLP hot path (SURVEY 2: RBFLayer out of scope)")
        if pos_enc:
            raise NotImplementedError('pos_enc encoder is not reachable from coords_split and is out of scope')
        self.image_downsampling = ImageDownsampling(sidelength=kwargs.get('sidelength', None),
                                                    downsample=kwargs.get('downsample', False))
        if not split_mlp:
            self.net = FCBlock(in_features=in_features, out_features=out_features, num_hidden_layers=num_hidden_layers,
                               hidden_features=hidden_features, outermost_linear=True, nonlinearity=type,
                               approx_layers=kwargs.get('approx_layers', 2), fusion_size=kwargs.get('fusion_size', 1),
                               reduced=kwargs.get('reduced', False))
            self.net.precision = precision
        else:
            self.net = SplitFCBlock(in_features=in_features, out_features=out_features,
                                    num_hidden_layers=num_hidden_layers, hidden_features=hidden_features,
                                    outermost_linear=True, nonlinearity=type, coord_dim=coord_dim, split_rule=split_rule,
                                    approx_layers=kwargs.get('approx_layers', 2), act_scale=kwargs.get('act_scale', 1),
                                    fusion_operator=kwargs.get('fusion_operator', 'prod'),
                                    fusion_before_act=kwargs.get('fusion_before_act', False),
                                    use_atten=kwargs.get('use_atten', False), learn_code=kwargs.get('learn_code', False),
                                    last_layer_features=kwargs.get('last_layer_features', -1),
                                    fusion_size=kwargs.get('fusion_size', 1), reduced=kwargs.get('reduced', False),
                                    precision=precision)
        if not quiet:
            print(self)

    def forward(self, model_input, params=None, ret_feat=False):
        """dict in / dict out (modules.py:419-454); if both keys are present the split result wins."""
        coords_org = output = None
        if 'coords' in model_input:
            coords_org = model_input['coords'].clone().detach().requires_grad_(True)
            coords = coords_org
            if self.mode == 'nerf':
                coords = self.positional_encoding(coords)
            output = self.net(coords)
        if 'coords_split' in model_input:
            coords_org = [c.clone().detach().requires_grad_(True) for c in model_input['coords_split']]
            coords = coords_org
            if self.mode == 'nerf':
                coords = [self.positional_encoding(c, single_channel=True) for c in coords]
            output = self.net(coords, split_coord=True, ret_feat=ret_feat)
        return {'model_in': coords_org, 'model_out': output}

    def forward_with_loss(self, model_input, gt, loss='mse', unit_grad=False):
        """forward() and the loss of the training step (training.py:112,133-135) in ONE call - an addition to the
        reference API: returns (model_output_dict, loss scalar).  loss: 'mse' = image_mse without mask
        (loss_functions.py:8-12), 'bce' = occupancy_3d (:56-60); gt: [1, P, O].  On the split grid path the loss is
        fused into the trunk's last layer (coordx_trunk_forward_loss: evaluated in the tcgen05 kernel's epilogue
        together with d loss / d model_out, no separate pass over the output); everything else takes forward()
        followed by loss_functions.fused_loss, with identical results."""
        from .loss_functions import fused_loss
        if loss not in ('mse', 'bce'):
            raise ValueError(f'unknown loss {loss!r}')
        hook = None
        if self.split_mlp and 'coords_split' in model_input and 'coords' not in model_input and gt.is_cuda:
            hook = {'kind': _lib.LOSS_MSE if loss == 'mse' else _lib.LOSS_BCE, 'gt': gt.detach(), 'loss': None, 'dout': None}
            self.net._loss_hook = hook
        try:
            output = self.forward(model_input)
        finally:
            if hook is not None:
                self.net._loss_hook = None
        out = output['model_out']
        if hook is not None and hook['loss'] is not None:
            value = _TrunkLossFn.apply(out, hook['loss'], hook['dout'], unit_grad) if out.requires_grad else hook['loss']
        else:
            value = fused_loss(loss, out, gt, unit_grad=unit_grad)
        return output, value

    def forward_split_channel(self, coord, channel_id):
        """modules.py:463-470"""
        if not self.split_mlp:
            return None
        if self.mode == 'nerf':
            coord = self.positional_encoding(coord, single_channel=True)
        return self.net.forward_channel(coord, channel_id)

    def forward_split_fusion(self, h):
        """modules.py:472-474"""
        return self.net.forward_fusion(h)
