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
