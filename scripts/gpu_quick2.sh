#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e pts/s  ms/step %.3f  e2e %.3e  kernel_ms %.4f  achieved %.1f TF  frac %.3f  launches/step %s clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['frac'], d['gpu_launches_per_step'], d['clocks']))
        print('train', d['train'])
    else: print(l.rstrip())
"
tail -n 5 gpurun_out/bench.err
