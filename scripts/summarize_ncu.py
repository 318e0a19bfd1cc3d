#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_fwd.csv  > profiles/r01_fwd_launches.txt
  python scripts/summarize_ncu.py full     gpurun_out/prof_fwd.ncu-rep  > profiles/r01_fwd_trunk_full.txt
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = re.compile(
    r'^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|sm__cycles_(elapsed|active)\.avg|'
    r'sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed|'
    r'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime\.avg|'
    r'sm__inst_executed_pipe_(xu|fma|alu|fmaheavy|fmalite|lsu|uniform)\.avg\.pct_of_peak_sustained_active|'
    r'sm__inst_executed_pipe_xu_realtime\.avg\.pct_of_peak_sustained_elapsed|'
    r'l1tex__data_pipe_(lsu|tc)_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|'
    r'l1tex__data_pipe_(lsu|tc)_wavefronts_mem_shared\.sum|'
    r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|'
    r'launch__(registers_per_thread|shared_mem_per_block_dynamic|grid_size|block_size|cluster.*)|'
    r'smsp__warp_issue_stalled_.*_per_warp_active\.pct|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
    r'sm__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|'
    r'sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|'
    r'lts__t_bytes\.sum|gpc__cycles_elapsed\.avg\.per_second)$')


def launches(path):
    rows = list(csv.reader(open(path, errors='replace')))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    iname, ival = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg, total, n = OrderedDict(), 0.0, 0
    for r in rows[start + 1:]:
        if len(r) <= ival:
            continue
        v = float(r[ival].replace(',', ''))
        a = agg.setdefault(r[iname], [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    print(f'# {path}: {n} launches, {total / 1e3:.1f} us of device time (ncu: cold-cache, serialised; compare SHARES)')
    print(f'{"count":>6} {"total_us":>11} {"share":>7} {"avg_us":>10}  kernel')
    for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{c:6d} {v / 1e3:11.1f} {100 * v / total:6.1f}% {v / c / 1e3:10.1f}  {name[:110]}')


def launches_multi(path):
    """launch list with several metrics per launch: one row per launch"""
    rows = list(csv.reader(open(path, errors='replace')))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    iid, iname, imet, ival, iunit = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    per = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= ival:
            continue
        d = per.setdefault(r[iid], {'kernel': r[iname]})
        d[r[imet]] = (float(r[ival].replace(',', '')), r[iunit])
    print(f'# {path}: one row per launch (ncu --clock-control none: cold-cache, serialised)')
    print(f'{"dur_us":>9} {"tensor%":>8} {"xu%":>6} {"issue%":>7} {"dram_rd_MB":>11} {"dram_wr_MB":>11} {"sm_GHz":>7}  kernel')
    for d in per.values():
        g = lambda k: d.get(k, (float('nan'), ''))[0]
        dur = g('gpu__time_duration.sum')
        dur_us = dur / 1e3 if d.get('gpu__time_duration.sum', (0, 'ns'))[1] in ('ns', 'nsecond') else dur
        print(f'{dur_us:9.1f} {g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):8.1f} '
              f'{g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"):6.1f} '
              f'{g("smsp__issue_active.avg.pct_of_peak_sustained_active"):7.1f} {g("dram__bytes_read.sum") / 1e6:11.1f} '
              f'{g("dram__bytes_write.sum") / 1e6:11.1f} {g("gpc__cycles_elapsed.avg.per_second") / 1e9:7.3f}  {d["kernel"][:90]}')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f'# {path}: ncu --set full --clock-control none; selected metrics per captured launch')
    for r in rows[2:]:
        print(f'\n== {r[hdr.index("Kernel Name")]}  grid {r[hdr.index("Grid Size")]} block {r[hdr.index("Block Size")]}')
        for h, u, v in zip(hdr, units, r):
            if KEEP.match(h):
                print(f'  {h:95s} {v:>16s} {u}')


if __name__ == '__main__':
    {'launches': launches, 'launches_multi': launches_multi, 'full': full}[sys.argv[1]](sys.argv[2])
