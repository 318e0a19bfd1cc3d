"""Launch-overhead tools around the split-MLP hot path.

`CapturedGrid` records one whole-grid inference - the reference's two-stage --speed_test sequence
(train_img.py:125-135: forward_split_channel per axis, then forward_split_fusion) - into a CUDA graph for
fixed axis lengths, so a step is ONE graph launch instead of a dozen Python-driven launches.  At 0.4 ms
of GPU work per 2048x2048 grid the Python gaps between launches are otherwise ~25 % of the step.

`infer_grid_host` is the host-buffer path of the C ABI (coordx_infer_grid_host): pinned host coordinates in,
pinned host result out, every copy and kernel enqueued by one C call.
"""
from __future__ import annotations

from typing import List, Sequence

import os

import torch

from . import _lib
from . import modules as M


class CapturedGrid:
    """CUDA-graph replay of model.forward_split_channel x C + model.forward_split_fusion for fixed shapes.

    axes: per-branch coordinate tensors [N_c, f_c] on the model's device.  `replay()` returns the static
    output tensor [N_0, ..., N_{C-1}, O]; `set_axes()` overwrites the static inputs in place."""

    def __init__(self, model: 'M.SingleBVPNet', axes: Sequence[torch.Tensor], warmup: int = 3):
        if not model.split_mlp:
            raise ValueError('CapturedGrid needs a split_mlp model')
        self.model = model
        self.static_axes = [a.detach().clone() for a in axes]
        C = len(self.static_axes)
        self.streams = [torch.cuda.Stream() for _ in range(C)]
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # first-launch work (function attributes, module load) stays outside
                self._forward(C)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self._forward(C)

    def _forward(self, C):
        # the C branch MLPs are independent: fork one stream per branch (parallel graph branches), join for the trunk
        main = torch.cuda.current_stream()
        feats = [None] * C
        for c, a in enumerate(self.static_axes):
            st = self.streams[c]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                h = self.model.forward_split_channel(a, c)
            shape = [1] * C + [h.shape[-1]]
            shape[c] = h.shape[0]
            feats[c] = h.reshape(shape)
        for st in self.streams:
            main.wait_stream(st)
        return self.model.forward_split_fusion(feats)

    def set_axes(self, axes: Sequence[torch.Tensor]) -> None:
        for s, a in zip(self.static_axes, axes):
            s.copy_(a, non_blocking=True)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.static_out


def infer_grid_host(model: 'M.SingleBVPNet', coords_host: Sequence[torch.Tensor], out_host: torch.Tensor,
                    precision: str = None, ws_cache: List[torch.Tensor] = None) -> torch.Tensor:
    """Whole-grid inference with HOST buffers through ONE C-ABI call (coordx_infer_grid_host).
    coords_host[c]: pinned fp32 [N_c, f_c] (already positionally encoded for nerf/fourier models);
    out_host: pinned fp32 with prod(N_c) * O elements.  Enqueues on the current stream; the caller
    synchronises before reading out_host."""
    lib = _lib.load()
    block = model.net
    params = block.abi_params()
    dev = params[0].device
    desc = M._net_desc(block, params)
    prec = block.resolve_precision(desc.net, precision)
    ns = [int(c.shape[0]) for c in coords_host]
    for c in coords_host:
        if c.is_cuda or c.dtype != torch.float32 or not c.is_contiguous():
            raise ValueError('coords_host must be contiguous fp32 host tensors')
    n_arr = _lib.i64_array(ns)
    c_arr = _lib.ptr_array([c.data_ptr() for c in coords_host])
    with torch.cuda.device(dev):
        ws_bytes = lib.coordx_infer_grid_host_workspace_bytes(desc.net, n_arr, prec)
        if ws_bytes == 0:
            raise _lib.CoordxError('coordx_infer_grid_host: unsupported configuration')
        ws = M._workspace(ws_bytes, dev)
        _lib.check(lib.coordx_infer_grid_host(desc.net, c_arr, n_arr, out_host.data_ptr(), prec, ws.data_ptr(),
                                              ws.numel(), torch.cuda.current_stream().cuda_stream), 'coordx_infer_grid_host')
    return out_host


class DeviceAdam:
    """torch.optim.Adam with its defaults (training.py:17), optionally preceded by clip_grad_norm_(max_grad_norm)
    (training.py:140-144) and followed by one ExponentialLR step (training.py:25,146), as ONE C call of three small
    launches (`coordx_optim_step`): the step count and the learning rate live on the device, so the whole optimiser
    step replays from a CUDA graph.  Differences from the torch objects it replaces: `p.grad` is read, not rescaled in
    place, by the clipping; state is `exp_avg` / `exp_avg_sq` lists plus `state` = [lr, steps, scratch, -]."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.0, lr_gamma=1.0):
        self.params = [p for p in params if p.requires_grad]
        if len(self.params) > 32:
            raise ValueError('DeviceAdam: at most 32 parameter tensors (COORDX_OPT_MAX_TENSORS)')
        if not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params):
            raise _lib.CoordxError('DeviceAdam: contiguous fp32 CUDA parameters required (no CPU path)')
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.state = torch.tensor([float(lr), 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        self.betas, self.eps, self.max_grad_norm, self.lr_gamma = betas, float(eps), float(max_grad_norm), float(lr_gamma)
        self._numel = _lib.i64_array([p.numel() for p in self.params])
        self._fixed = [_lib.ptr_array([t.data_ptr() for t in ts]) for ts in (self.params, self.exp_avg, self.exp_avg_sq)]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self):
        ps = [(i, p) for i, p in enumerate(self.params) if p.grad is not None]
        if not ps:
            return
        if len(ps) == len(self.params):
            pp, mm, vv, nn = self._fixed[0], self._fixed[1], self._fixed[2], self._numel
        else:                                        # a parameter without gradient this step is skipped, as torch does
            pp = _lib.ptr_array([p.data_ptr() for _, p in ps])
            mm = _lib.ptr_array([self.exp_avg[i].data_ptr() for i, _ in ps])
            vv = _lib.ptr_array([self.exp_avg_sq[i].data_ptr() for i, _ in ps])
            nn = _lib.i64_array([p.numel() for _, p in ps])
        for _, p in ps:
            if not (p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                p.grad = p.grad.contiguous().float()
        gg = _lib.ptr_array([p.grad.data_ptr() for _, p in ps])
        dev = self.state.device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().coordx_optim_step(len(ps), pp, gg, mm, vv, nn, self.state.data_ptr(), self.betas[0],
                                                     self.betas[1], self.eps, self.max_grad_norm, self.lr_gamma,
                                                     torch.cuda.current_stream(dev).cuda_stream), 'coordx_optim_step')

    def snapshot(self):
        return [t.clone() for t in (*self.exp_avg, *self.exp_avg_sq, self.state)]

    def restore(self, snap):
        with torch.no_grad():
            for t, s in zip((*self.exp_avg, *self.exp_avg_sq, self.state), snap):
                t.copy_(s)


class CapturedTrainStep:
    """CUDA-graph replay of ONE training iteration of training.py:112-146 for fixed shapes:
    model({'coords_split': ...}) -> loss -> backward -> [clip_grad_norm_ 1.0] -> Adam.step, i.e. the body of the
    reference's epoch loop without its TensorBoard / .item() synchronisations.  A step is one graph launch instead
    of ~130 Python-driven launches: the small named shapes (256x256 image, 32^3 occupancy batches) are otherwise
    bound by launch overhead, not by the GPU.

    coords: the 'coords_split' list ([1, N_0, 1, .., f_0], ...); gt: [1, P, O].  loss: 'mse' (loss_functions.py:8-12),
    'bce' (:56-60) or a callable(model_out, gt) -> scalar.  `replay()` returns the static loss tensor (device);
    `set_batch()` overwrites the static inputs in place.  The warm-up iterations needed before capture are undone:
    parameters and optimizer state are restored to their values at construction."""

    def __init__(self, model: 'M.SingleBVPNet', coords: Sequence[torch.Tensor], gt: torch.Tensor, loss='mse',
                 lr: float = 1e-4, clip_grad: bool = False, optimizer: torch.optim.Optimizer = None, warmup: int = 3,
                 lr_gamma: float = None, world: int = 1, group=None, grad_weight: float = None):
        self.model = model
        # world > 1: the grid is sharded along axis 0 (sharding.py) and `coords` / `gt` are this rank's slab; the one
        # exchange of the step - the flat all-reduce of the weight gradients - is captured in the graph with the rest
        # (NCCL collectives are capturable), so a multi-GPU step is still ONE graph launch per rank
        # grad_weight: this rank's share of the grid points (sharding.slab_weight) when the slabs are not equally long
        self.world, self.group, self.grad_weight = int(world), group, grad_weight
        self.static_coords = [c.detach().clone() for c in coords]
        self.static_gt = gt.detach().clone()
        self.clip_grad = clip_grad
        self.params = [p for p in model.parameters() if p.requires_grad]
        # Default optimiser: DeviceAdam - clip_grad_norm_ (if clip_grad), Adam and the reference's ExponentialLR
        # (lr_gamma; training.py:25, stepped once per iteration :146) in one C call with device-side step / lr state.
        # A caller-supplied torch optimiser must be capturable (and cannot take lr_gamma).
        self.lr_gamma = lr_gamma
        if optimizer is None:
            self.opt = DeviceAdam(self.params, lr=lr, max_grad_norm=1.0 if clip_grad else 0.0,
                                  lr_gamma=1.0 if lr_gamma is None else lr_gamma)
            self._own_opt = True
        else:
            if lr_gamma is not None:
                raise ValueError('lr_gamma needs the built-in optimizer (its learning rate is device state)')
            if not all(g.get('capturable', False) for g in optimizer.param_groups):
                raise ValueError('CapturedTrainStep needs an optimizer constructed with capturable=True')
            self.opt = optimizer
            self._own_opt = False
        # COORDX_FUSED_LOSS=0: forward + separate loss kernel (A/B timing of the loss fused into the trunk)
        self.loss_kind = loss if loss in ('mse', 'bce') and os.environ.get('COORDX_FUSED_LOSS', '1') != '0' else None
        if callable(loss):
            self.loss_fn = loss
        elif loss in ('mse', 'bce'):                 # one kernel: value and d loss / d out (loss_functions.py)
            from .loss_functions import fused_loss
            self.loss_fn = lambda out, gt_: fused_loss(loss, out, gt_, unit_grad=True)
        else:
            raise ValueError(f'unknown loss {loss!r}')
        saved_p = [p.detach().clone() for p in self.params]
        if self._own_opt:
            saved_s = self.opt.snapshot()
        else:
            saved_s = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.opt.state.get(p, {}).items()}
                       for p in self.params}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self.opt.zero_grad(set_to_none=True)
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._iteration()
        with torch.no_grad():                                    # undo the warm-up updates
            for p, s in zip(self.params, saved_p):
                p.copy_(s)
            if self._own_opt:
                self.opt.restore(saved_s)
            else:
                for p in self.params:
                    st, old = self.opt.state.get(p, {}), saved_s[id(p)]
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.copy_(old[k]) if k in old else v.zero_()
        torch.cuda.synchronize()

    def _iteration(self):
        if self.loss_kind is not None and hasattr(self.model, 'forward_with_loss'):
            # loss fused into the trunk's last layer (coordx_trunk_forward_loss)
            _, loss = self.model.forward_with_loss({'coords_split': self.static_coords}, self.static_gt, self.loss_kind,
                                                   unit_grad=True)
        else:
            out = self.model({'coords_split': self.static_coords})['model_out']
            loss = self.loss_fn(out, self.static_gt)
        loss.backward()
        if self.world > 1:
            from . import sharding
            sharding.allreduce_grads(self.params, self.world, group=self.group, weight=self.grad_weight)
        if self.clip_grad and not self._own_opt:                 # (DeviceAdam clips inside its step)
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=1.)
        self.opt.step()
        return loss

    def set_batch(self, coords: Sequence[torch.Tensor] = None, gt: torch.Tensor = None) -> None:
        if coords is not None:
            for s, c in zip(self.static_coords, coords):
                s.copy_(c, non_blocking=True)
        if gt is not None:
            self.static_gt.copy_(gt, non_blocking=True)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.static_loss
