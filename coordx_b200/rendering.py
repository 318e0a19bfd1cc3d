"""Feature-interpolated evaluation of scattered points from CACHED per-axis branch features - the trunk side of the
reference's split volume rendering (ray_rendering.py:133-145 `get_pts_pred(split=True)`, called per ray-sample batch by
`vol_render_split`, :196-295).

The branch MLPs are evaluated once per axis on a regular grid (`model.forward_split_channel`); a query point then only
needs a gather + linear interpolation of two rows per axis, the fusion, and the trunk.  Here the gather, the lerp and
the fusion are ONE kernel (`coordx_interp_fuse`: the tables stay in L2, only the fused [P, Hd] features are written),
and the trunk is the same tcgen05 / separable kernel as on the grid (a 1 x P grid whose first axis is the fusion's
neutral element).  Inference only, like the reference (`torch.no_grad()` around the trunk, ray_rendering.py:143-144)."""
import torch

from . import _lib
from . import modules as M


def interp_fuse(block: 'M.SplitFCBlock', pts_idx: torch.Tensor, feats) -> torch.Tensor:
    """pts_idx [..., C] fractional indices into the C feature tables feats[c] [n_c, F]  ->  fused features [P, F]."""
    lib = _lib.load()
    C = len(feats)
    if C < 1 or C > _lib.MAX_BRANCHES or pts_idx.shape[-1] != C:
        raise ValueError(f'interp_fuse: {C} feature tables for indices of width {pts_idx.shape[-1]}')
    tables = [M._as_f32c(f.reshape(-1, f.shape[-1])) for f in feats]
    idx = M._as_f32c(pts_idx.reshape(-1, C))
    if not (idx.is_cuda and all(t.is_cuda for t in tables)):
        raise _lib.CoordxError('coordx_b200.rendering: CUDA tensors required (no CPU path)')
    F = tables[0].shape[-1]
    if F % 4 or any(t.shape[-1] != F for t in tables):
        raise ValueError('interp_fuse: feature tables must share a width that is a multiple of 4')
    P = idx.shape[0]
    h = torch.empty(P, F, dtype=torch.float32, device=idx.device)
    if P == 0:
        return h
    op = {'sum': _lib.FUSE_SUM, 'prod': _lib.FUSE_PROD}[block.fusion_operator]
    with torch.cuda.device(idx.device):
        _lib.check(lib.coordx_interp_fuse(C, _lib.ptr_array([t.data_ptr() for t in tables]),
                                          _lib.i64_array([t.shape[0] for t in tables]), F, op, idx.data_ptr(), P, h.data_ptr(),
                                          torch.cuda.current_stream(idx.device).cuda_stream), 'coordx_interp_fuse')
    return h


def get_pts_pred(model: 'M.SingleBVPNet', pts_idx: torch.Tensor, feats, split: bool = True, resolution=None) -> torch.Tensor:
    """ray_rendering.py:133-145 with split=True: pts_idx [P, C] (fractional grid indices), feats = per-axis feature
    tables from `model.forward_split_channel` on the grid lines -> model output [P, out_features]."""
    if not split:
        raise NotImplementedError('get_pts_pred(split=False) interpolates a dense [R^3, O] prediction volume: no network '
                                  'evaluation on that path (ray_rendering.py:146-170)')
    block = model.net
    with torch.no_grad():
        h = interp_fuse(block, pts_idx, feats)
        if h.shape[0] == 0:
            return h.new_zeros(*pts_idx.shape[:-1], block.out_width)
        neutral = (torch.zeros if block.fusion_operator == 'sum' else torch.ones)(1, h.shape[-1], dtype=h.dtype, device=h.device)
        out = M._fusion_forward(block, None, [neutral, h], block.abi_params())[0]
    return out.reshape(*pts_idx.shape[:-1], out.shape[-1])
