#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "tensor_core or tc_" 2>&1 | tail -n 3
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e pts/s  ms/step %.3f  e2e %.3e  kernel_ms %.4f  achieved %.1f TF  frac %.3f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['frac'], d['clocks']))
    else: print(l.rstrip())
"
