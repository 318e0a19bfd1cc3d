#!/bin/bash
# forward-path check: tcgen05 parity tests, then a short forward-only bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core_forward or tc_large or two_stage" > gpurun_out/t_fwd.log 2>&1; echo "tests exit $?"; tail -n 15 gpurun_out/t_fwd.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-cpu 2>gpurun_out/bench_fwd.err | tee gpurun_out/bench_fwd.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e pts/s  ms/step %.3f  e2e %.3e  kernel_ms %.4f  achieved %.1f TF  frac %.3f  launches/step %s clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['frac'], d['gpu_launches_per_step'], d['clocks']))
    else: print(l.rstrip())
"
tail -n 5 gpurun_out/bench_fwd.err
