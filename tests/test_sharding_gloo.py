"""N>1 host logic on CPU: world_size-2 `gloo` processes shard the grid along axis 0,
each computes the gradient of its slab (numpy oracle as the per-rank compute), and ONE flat
all-reduce reproduces the full-grid gradient.  No GPU, no product kernels: this covers
coordx_b200/sharding.py (slab_range / shard_coords / shard_target / allreduce_grads / gather_output)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coordx_b200 import sharding
from oracle import splitmlp_np as O


def test_slab_range_partitions():
    for n0 in (1, 7, 16, 2048, 2160):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.slab_range(n0, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n0
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.slab_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, H=10):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cfg = O.SplitCfg(in_features=2, coord_dim=2, out_features=1, hidden_features=32, fusion_operator='prod')
        p = O.init_params(cfg, np.random.default_rng(0), dtype=np.float64)
        W = 6                                          # H = 10 rows over 2 ranks: 5 + 5; H = 9: 5 + 4
        xs = torch.linspace(-1, 1, H, dtype=torch.float64).reshape(1, H, 1, 1)
        ys = torch.linspace(-1, 1, W, dtype=torch.float64).reshape(1, 1, W, 1)
        gt = torch.as_tensor(np.random.default_rng(1).uniform(-1, 1, (1, H * W, 1)))
        mine = sharding.shard_coords([xs, ys], rank, world)
        gt_mine = sharding.shard_target(gt, (H, W), rank, world)
        coords = [c.reshape(-1, 1).numpy() for c in mine]
        cache = {}
        out, _ = O.forward_grid(p, cfg, coords, cache)
        loss, dout = O.image_mse(out, gt_mine[0].numpy())     # mean over this rank's slab
        grads, _ = O.backward_grid(p, cfg, cache, dout)
        params = []
        for k in sorted(grads):
            t = torch.nn.Parameter(torch.as_tensor(p[k].copy()))
            t.grad = torch.as_tensor(np.ascontiguousarray(grads[k]))
            params.append(t)
        if H % world == 0:
            sharding.allreduce_grads(params, world, average=True)          # equal slabs: plain mean
        else:
            sharding.allreduce_grads(params, world, weight=sharding.slab_weight(H, rank, world))
        full = sharding.gather_output(torch.as_tensor(out)[None], world)
        if rank == 0:
            # single-process truth on the whole grid
            cache = {}
            out_f, _ = O.forward_grid(p, cfg, [xs.reshape(-1, 1).numpy(), ys.reshape(-1, 1).numpy()], cache)
            _, dout_f = O.image_mse(out_f, gt[0].numpy())
            g_f, _ = O.backward_grid(p, cfg, cache, dout_f)
            err = max(float(np.abs(t.grad.numpy() - g_f[k]).max() / (np.abs(g_f[k]).max() + 1e-30))
                      for k, t in zip(sorted(grads), params))
            q.put(('ok', err, float(np.abs(full[0].numpy() - out_f).max())))
    except Exception as e:                              # pragma: no cover
        if rank == 0:
            q.put(('err', repr(e), 0.0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('H', [10, 9])
def test_two_rank_gloo_grad_allreduce_equals_full_grid(H):
    """H = 9: unequal slabs (5 + 4 rows) - gradients weighted by the slab's share, outputs gathered with padding."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, H)) for r in range(2)]
    for pr in procs:
        pr.start()
    status, gerr, oerr = q.get(timeout=180)
    for pr in procs:
        pr.join(60)
    assert status == 'ok', gerr
    assert gerr < 1e-10 and oerr < 1e-12
