"""Multi-GPU sharding of the split-MLP grid.

Grid points are independent given the weights, so a grid shards along ONE split axis
(axis 0: image rows / video frames / occupancy x - the axis the reference already loops
over for memory, train_3d_occupancy.py:142-144) with no forward communication: every rank
evaluates the branch MLPs it needs (its slab of axis 0, all of the other axes) and its slab
of the trunk.  Training adds exactly one exchange: the weight-gradient all-reduce (0.8 MB for
the 256-wide nets) as ONE flat NCCL all-reduce, after which every rank applies the same Adam
step (training.py:146).  The reference itself has no distributed code.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def slab_range(n0: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slab of axis 0 for `rank` (first n0 % world ranks get one more row)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank/world {rank}/{world}')
    base, extra = divmod(int(n0), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_coords(coords_split: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    """Slice the axis-0 coordinate tensor ([B, N_0, 1, .., f_0]) to this rank's slab; other axes stay whole."""
    a, b = slab_range(coords_split[0].shape[1], rank, world)
    return [coords_split[0][:, a:b]] + list(coords_split[1:])


def shard_target(gt: torch.Tensor, shape: Sequence[int], rank: int, world: int) -> torch.Tensor:
    """gt [B, P, O] flattened row-major over `shape` -> the rows of this rank's slab."""
    inner = 1
    for n in shape[1:]:
        inner *= int(n)
    a, b = slab_range(shape[0], rank, world)
    return gt[:, a * inner:b * inner]


def allreduce_grads(params: Iterable[torch.nn.Parameter], world: int, average: bool = True, group=None,
                    weight: float | None = None) -> None:
    """One flat all-reduce of every parameter gradient.

    average=True, weight=None: plain mean over ranks - the gradient of the global mean loss when every rank's loss is
    a mean over an EQUALLY sized slab.  When the slabs differ (n0 % world != 0, `slab_range` hands out unequal
    slabs), pass `weight` = this rank's share of the points (`slab_weight`): the per-rank mean-loss gradients are
    then combined as sum_r weight_r * grad_r, which is the global-mean gradient for any split."""
    if world <= 1:
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    if weight is not None:
        flat *= float(weight)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average and weight is None:
        flat /= world
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def slab_weight(n0: int, rank: int, world: int) -> float:
    """This rank's share of the grid points (its slab's rows over all rows): the `weight` of `allreduce_grads`."""
    a, b = slab_range(n0, rank, world)
    return (b - a) / float(n0) if n0 > 0 else 0.0


def gather_output(out_slab: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """Optional: assemble the full [B, P, O] output from the ranks' slabs (inference can also stay sharded).
    Slabs may differ in length along dim 1 (n0 % world != 0): sizes are exchanged first and the slabs padded to the
    longest for the all_gather."""
    if world <= 1:
        return out_slab
    out_slab = out_slab.contiguous()
    mine = torch.tensor([out_slab.shape[1]], dtype=torch.int64, device=out_slab.device)
    sizes = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    sizes = [int(t.item()) for t in sizes]
    longest = max(sizes)
    if out_slab.shape[1] < longest:
        pad = out_slab.new_zeros(out_slab.shape[0], longest - out_slab.shape[1], *out_slab.shape[2:])
        out_slab = torch.cat([out_slab, pad], dim=1)
    parts = [torch.empty_like(out_slab) for _ in range(world)]
    dist.all_gather(parts, out_slab, group=group)
    return torch.cat([p[:, :k] for p, k in zip(parts, sizes)], dim=1)
