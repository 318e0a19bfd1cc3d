#!/usr/bin/env python
"""bench.py - headline benchmark of the split-MLP hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one synthetic grid: both stages of the
reference's --speed_test loop (train_img.py:125-135): per-axis branch MLPs
(forward_split_channel x C) + fusion and trunk over every grid point
(forward_split_fusion).  Workload at N=1 = BASELINE.json configs[1]: 2048x2048 image,
sine, fusion=prod, hidden 256, bf16 trunk.  At N>1 every rank owns a 2048-row slab of a
(2048*N) x 2048 grid (weak scaling, no data-path collective).

Prints ONE JSON line (see the task contract): value = grid points/s with inputs resident
in HBM; e2e = the same through the public python API with pinned-host coordinates in and
the [P,1] result copied back to pinned host memory each step; roofline for the dominant
(tcgen05 trunk) kernel; cpu_baseline = the oracle port (torch CPU eager, the reference's own
op sequence) on a bounded sample; train = one training iteration (fwd + MSE + bwd + Adam).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 2048
HIDDEN = 256


def flops_fwd(n_axes, hidden=HIDDEN, out=1, al=2, nhl=3, f=1):
    """F_fwd = P(2(nhl-al)Hd^2 + 2 Hd O) + sum_c N_c (2 f_c Hd + 2 al Hd^2)"""
    P = 1
    for n in n_axes:
        P *= n
    return P * (2 * (nhl - al) * hidden ** 2 + 2 * hidden * out) + sum(n * (2 * f * hidden + 2 * al * hidden ** 2) for n in n_axes)


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            p = json.load(fh)
        return p, 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, smax, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = max(smax, float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        # every retained sample was taken during the sustained probe loop (the rows are cleared right before it)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_min_mhz': min(sm) if sm else None,
                'sm_max_mhz': smax or None, 'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------- reference arm
def cpu_reference(steps, warmup, budget_s=20.0, train=False):
    """The reference's own CPU implementation of the path = torch CPU eager with the reference's op
    sequence (oracle/ref_port_torch.py, validated against tests/golden).  Bounded sample: the
    2048x2048 grid is evaluated in 128-row slabs along axis 0 (train_3d_occupancy.py:142-144);
    each step times `slab_limit` slabs."""
    import torch
    from oracle import ref_port_torch as RP, splitmlp_np as O
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.SplitCfg(in_features=2, coord_dim=2, out_features=1, hidden_features=HIDDEN, fusion_operator='prod')
    p = RP.to_torch_params(O.init_params(cfg, np.random.default_rng(0)))
    axes = [torch.linspace(-1, 1, H)[:, None], torch.linspace(-1, 1, W)[:, None]]
    slabs, slab_rows = 16, H // 16
    # calibrate: one slab
    t0 = time.perf_counter()
    RP.infer_grid(p, cfg, axes, slabs=slabs, slab_limit=1)
    t1 = time.perf_counter() - t0
    per_step_budget = budget_s / max(steps + warmup, 1)
    k = int(max(1, min(slabs, per_step_budget / max(t1, 1e-3))))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        RP.infer_grid(p, cfg, axes, slabs=slabs, slab_limit=k)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    pts = k * slab_rows * W
    ms = 1e3 * sum(times) / len(times)
    return {'value': pts / (ms * 1e-3), 'ms_per_step': ms, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{k} of {slabs} axis-0 slabs ({k * slab_rows}x{W} points) of the 2048x2048 grid per step, '
                      f'branch MLPs for all {H}+{W} axis samples included; torch {torch.__version__} CPU eager'}


def main():
    # stdout carries exactly ONE line (the JSON): everything else that libraries print there (e.g. NCCL's version
    # banner at communicator creation) is sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--train-steps', type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    warmup = max(args.warmup, 3)
    metric, unit = 'grid points/sec (fwd)', 'points/s'
    config = {'workload': f'train_img.py synthetic {H}x{W} image per GPU, sine split MLP, fusion=prod, hidden 256, '
                          f'{args.precision} trunk (BASELINE configs[1])',
              'grid_per_gpu': [H, W], 'hidden': HIDDEN, 'fusion': 'prod', 'l2': 'flushed between timed iterations (256 MiB write)',
              'launch': 'value: CUDA-graph replay of forward_split_channel x2 + forward_split_fusion; e2e: one coordx_infer_grid_host call',
              'sharding': 'axis-0 slab per rank, no data-path collective'}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_reference(args.steps, warmup)
        line = {'impl': 'reference', 'metric': metric, 'value': r['value'], 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps,
                'warmup': warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': r['value'], 'unit': unit, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
                'e2e': {'value': r['value'], 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import ctypes
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__ as G
    if rank == 0:
        G.build()
    if world > 1:
        dist.barrier()
    from coordx_b200 import modules as M, _lib
    lib = _lib.load()

    torch.manual_seed(0)
    model = M.SingleBVPNet(type='sine', mode='mlp', out_features=1, hidden_features=HIDDEN, split_mlp=True,
                           fusion_operator='prod', quiet=True, precision=args.precision).to(dev)
    # this rank's slab of the (H*world) x W grid
    rows_total = H * world
    x_all = torch.linspace(-1, 1, rows_total)[:, None]
    x_host = x_all[rank * H:(rank + 1) * H].contiguous().pin_memory()
    y_host = torch.linspace(-1, 1, W)[:, None].contiguous().pin_memory()
    x, y = x_host.to(dev), y_host.to(dev)
    out_host = torch.empty(H * W, 1).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    P = H * W

    def fwd(xd, yd):
        fx_ = model.forward_split_channel(xd, 0).unsqueeze(1)
        fy_ = model.forward_split_channel(yd, 1).unsqueeze(0)
        return model.forward_split_fusion([fx_, fy_])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        """W untimed + K timed iterations; L2 flushed before each timed iteration; device time
        from CUDA events on the launching stream, summed over the K iterations."""
        for _ in range(warm):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        per = [a.elapsed_time(b) for a, b in evs]
        ms = sum(per)
        if max(per) > 3 * statistics.median(per):            # diagnostics only (stderr): an outlier inside the timed region
            print(f'[bench rank {rank}] uneven steps (ms): ' + ' '.join(f'{v:.2f}' for v in per), file=sys.stderr, flush=True)
        med = statistics.median(per) * len(per)
        if world > 1:
            t = torch.tensor([ms, med], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, med = float(t[0].item()), float(t[1].item())
        timed.last_median_total = med                         # K x the median step (max over ranks): robust to one stalled step
        return ms

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)

    # ---- value: inputs resident in HBM ---------------------------------------------------
    # One step = the reference's two-stage --speed_test sequence (train_img.py:125-135).  It is replayed from a
    # CUDA graph (coordx_b200.runtime.CapturedGrid): ~0.4 ms of GPU work per step is otherwise stretched ~25 % by
    # Python gaps between the launches.  The eager (no graph) figure is reported next to it.
    from coordx_b200 import runtime as RT
    kernel_ms = []
    with torch.no_grad():
        launches0 = lib.coordx_launch_count()
        fwd(x, y)
        launches_per_step = lib.coordx_launch_count() - launches0
        cap = RT.CapturedGrid(model, [x, y])
        total_ms = timed(cap.replay, args.steps, warmup)
        eager_ms = timed(lambda: fwd(x, y), args.steps, warmup)
        # dominant kernel alone: a few extra eager launches, each read back from its own event pair
        lib.coordx_profile_enable(1)
        for _ in range(5):
            flush.zero_()
            fwd(x, y)
            v = ctypes.c_float(0)
            _lib.check(lib.coordx_profile_last_ms(ctypes.byref(v)), 'profile')
            kernel_ms.append(v.value)
        lib.coordx_profile_enable(0)

        # ---- sustained probe: the timed region above is ~10 ms, far too short for nvidia-smi to sample clocks under
        # load; the same graph is replayed back to back for ~1.5 s (no L2 flush) while the sampler runs
        sampler.rows.clear()
        barrier()
        t_probe0 = time.perf_counter()
        n_probe = 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        while time.perf_counter() - t_probe0 < 1.5:
            for _ in range(50):
                cap.replay()
            n_probe += 50
            torch.cuda.synchronize()
        ev1.record()
        torch.cuda.synchronize()
        sustained_ms = ev0.elapsed_time(ev1) / max(n_probe, 1)
        clocks = sampler.finish()

        # ---- e2e: pinned host coordinates in, [P,1] result to pinned host memory, ONE C-ABI call
        # (coordx_infer_grid_host: H2D copies, branch MLPs, trunk, D2H copy all inside the timed region)
        def step_e2e():
            RT.infer_grid_host(model, [x_host, y_host], out_host)
        e2e_ms = timed(step_e2e, args.steps, warmup)
        # the result that came back over PCIe is the same as the graph's
        torch.cuda.synchronize()
        e2e_check = float((out_host.reshape(-1) - cap.replay().reshape(-1).cpu()).abs().max())

    ms_per_step = total_ms / args.steps
    value = P * world / (ms_per_step * 1e-3)
    e2e_value = P * world / (e2e_ms / args.steps * 1e-3)
    pk, pk_kind = peaks()
    k_ms = statistics.mean(kernel_ms)
    F = flops_fwd([H, W])
    F_kernel = P * (2 * HIDDEN * HIDDEN + 2 * HIDDEN)          # trunk kernel's share of F_fwd (SURVEY 8d per-point 131584)
    achieved = F_kernel / (k_ms * 1e-3) / 1e12
    roof = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
            'frac': achieved / pk['bf16_tflops'], 'traffic': None, 'peak_kind': pk_kind + ' burst (kernel timed alone)',
            'kernel': 'trunk_fwd_tc_kernel', 'kernel_ms': k_ms, 'flop_per_launch': F_kernel,
            'step_frac_of_sustained': F / (ms_per_step * 1e-3) / 1e12 / pk.get('bf16_tflops_sustained', pk['bf16_tflops']),
            'eager_ms_per_step': eager_ms / args.steps, 'e2e_vs_graph_max_abs_diff': e2e_check,
            'sustained_ms_per_step': sustained_ms, 'sustained_value': P * world / (sustained_ms * 1e-3),
            'sustained_note': 'graph replayed back to back for ~1.5 s (this rank), no L2 flush: the clocks entry is sampled during this loop'}
    # DRAM bytes of the trunk kernel per launch, from the committed ncu --set full capture (profiles/fwd_trunk_dram.json)
    try:
        with open(os.path.join(ROOT, 'profiles', 'fwd_trunk_dram.json')) as fh:
            tr = json.load(fh)
        roof['traffic'] = tr['dram_bytes_read'] + tr['dram_bytes_write']
        roof['traffic_note'] = tr['note']
    except Exception:
        pass

    # ---- training iteration (fwd + MSE + bwd + Adam), extra keys -------------------------------
    train = None
    if not args.no_train:
        gt = (torch.rand(1, P, 1, device=dev) * 2 - 1)
        gt_host = gt.cpu().pin_memory()
        opt = torch.optim.Adam(lr=1e-4, params=model.parameters(), fused=True)     # training.py:17 (fused = implementation detail)
        xs, ys = x.reshape(1, H, 1, 1), y.reshape(1, 1, W, 1)
        from coordx_b200 import loss_functions as LF

        def train_step(gtd=gt):
            mo = model({'coords_split': [xs, ys]})
            loss = LF.image_mse(None, mo, {'img': gtd})['img_loss']          # loss_functions.py:8-12 (drop-in module)
            opt.zero_grad()
            loss.backward()
            if world > 1:
                from coordx_b200 import sharding
                sharding.allreduce_grads(model.parameters(), world)
            opt.step()
            return loss
        if world > 1:                                         # NCCL sets its channels up lazily: not inside the timed region
            from coordx_b200 import sharding
            warm = torch.zeros(sum(p.numel() for p in model.parameters()), device=dev)
            for _ in range(3):
                dist.all_reduce(warm)
        tr_ms_sum = timed(train_step, args.train_steps, 5)
        tr_ms = timed.last_median_total
        loss_host = torch.empty(1).pin_memory()

        def train_e2e():
            gtd = gt_host.to(dev, non_blocking=True)
            l = train_step(gtd)
            loss_host.copy_(l.detach().reshape(1), non_blocking=True)
        tr_e2e_ms = timed(train_e2e, args.train_steps, 1)
        eager_it_s = args.train_steps / (tr_ms * 1e-3)
        best_ms = tr_ms / args.train_steps
        graph_it_s = None
        try:
            # the same iteration (fwd + MSE + bwd + [NCCL all-reduce of the weight gradients] + Adam) replayed from ONE
            # CUDA graph per rank (runtime.CapturedTrainStep)
            from coordx_b200.runtime import CapturedTrainStep
            cap = CapturedTrainStep(model, [xs, ys], gt, loss='mse', lr=1e-4, world=world)
            timed(cap.replay, args.train_steps, 3)
            g_ms = timed.last_median_total / args.train_steps
            graph_it_s = 1e3 / g_ms
            best_ms = min(best_ms, g_ms)
            del cap
        except Exception as e:                                 # the eager figure stands
            print(f'[bench rank {rank}] captured training step unavailable: {e!r}', file=sys.stderr, flush=True)
        train = {'metric': 'train iters/sec', 'value': 1e3 / best_ms, 'ms_per_iter': best_ms,
                 'eager_value': eager_it_s, 'graph_value': graph_it_s, 'eager_mean_ms_per_iter': tr_ms_sum / args.train_steps,
                 'timing': 'per-iteration CUDA events, MEDIAN iteration x count (max over ranks); eager_mean_ms_per_iter is the plain mean',
                 'e2e_value': args.train_steps / (tr_e2e_ms * 1e-3),
                 'tflops_algorithmic': 3 * F / (best_ms * 1e-3) / 1e12,
                 'frac_of_sustained_peak': 3 * F / (best_ms * 1e-3) / 1e12 / pk.get('bf16_tflops_sustained', pk['bf16_tflops']),
                 'note': 'one training iteration through the reference-shaped API: model(dict) -> MSE -> loss.backward() -> Adam; '
                         'forward and backward trunk on tcgen05 (bf16); eager_value = Python-driven launches, graph_value = the same '
                         'iteration replayed from one CUDA graph per rank (the NCCL all-reduce captured with it); global grid = this rank\'s slab, grads all-reduced over NCCL when N>1'}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference(3, 1, budget_s=15.0)
        cpu = {'value': r['value'], 'unit': unit, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}

    if rank == 0:
        line = {'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
                'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': args.precision, 'data': 'synthetic', 'config': config,
                'e2e': {'value': e2e_value, 'unit': unit, 'h2d_bytes_per_step': (H + W) * 4, 'd2h_bytes_per_step': P * 4},
                'gpu_launches': int(launches_per_step) * args.steps, 'gpu_launches_per_step': int(launches_per_step),
                'roofline': roof, 'cpu_baseline': cpu, 'clocks': clocks, 'train': train}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
