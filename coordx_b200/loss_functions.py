"""Losses of the split-MLP path with the reference's signatures (loss_functions.py:8-12 `image_mse`,
:56-60 `occupancy_3d`), evaluated by ONE kernel (`coordx_loss_forward_backward`): the loss value and
d loss / d model_out come out of the same pass over `model_out` / `gt`, so the backward of the loss is a saved
tensor instead of the reference's six elementwise kernels (sub, pow, mean, and their three backward kernels).

`fused_loss(kind, out, gt)` is the tensor-level entry (`kind` 'mse' | 'bce'); CUDA fp32 tensors only - like the rest
of the package there is no CPU path (the reference's own loss functions are plain torch and work anywhere)."""
import torch
from torch.autograd.function import once_differentiable

from . import _lib

_KINDS = {'mse': _lib.LOSS_MSE, 'bce': _lib.LOSS_BCE}


class _FusedLossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, kind, out, gt, unit_grad):
        lib = _lib.load()
        if not (out.is_cuda and gt.is_cuda):
            raise _lib.CoordxError('coordx_b200.loss_functions: CUDA tensors required (no CPU path)')
        o = out.detach().contiguous().float()
        g = gt.detach().expand_as(out).contiguous().float()
        loss = torch.zeros((), dtype=torch.float32, device=out.device)
        dout = torch.empty_like(o) if ctx.needs_input_grad[1] else None
        numel = o.numel()
        with torch.cuda.device(out.device):
            _lib.check(lib.coordx_loss_forward_backward(_KINDS[kind], o.data_ptr(), g.data_ptr(), numel, 1.0 / max(numel, 1),
                                                        loss.data_ptr(), dout.data_ptr() if dout is not None else None,
                                                        torch.cuda.current_stream(out.device).cuda_stream),
                       'coordx_loss_forward_backward')
        ctx.dout, ctx.unit_grad, ctx.shape = dout, unit_grad, out.shape
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        d = ctx.dout
        if d is None:
            return None, None, None, None
        if not ctx.unit_grad:                        # loss.backward() itself passes exactly 1
            d = d * g
        return None, d.reshape(ctx.shape), None, None


def fused_loss(kind, out, gt, unit_grad=False):
    """mean squared error ('mse') or binary cross-entropy with logits ('bce') over all elements.
    unit_grad=True promises that the result is differentiated with an upstream gradient of exactly 1
    (`loss.backward()` on this scalar): the saved gradient is returned as is, without a scaling kernel."""
    if kind not in _KINDS:
        raise ValueError(f'unknown loss {kind!r}')
    return _FusedLossFn.apply(kind, out, gt, unit_grad)


def image_mse(mask, model_output, gt):
    """loss_functions.py:8-12.  A mask takes the reference's elementwise expression."""
    if mask is None:
        return {'img_loss': fused_loss('mse', model_output['model_out'], gt['img'])}
    return {'img_loss': (mask * (model_output['model_out'] - gt['img']) ** 2).mean()}


def occupancy_3d(model_output, gt):
    """loss_functions.py:56-60."""
    return {'entropy': fused_loss('bce', model_output['model_out'], gt['occupancy'])}
