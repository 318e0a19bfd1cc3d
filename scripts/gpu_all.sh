#!/bin/bash
# full GPU check: every gpu-marked test, smoke, then the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 "$@" 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e pts/s  ms/step %.3f  e2e %.3e  kernel_ms %.4f  achieved %.1f TF  frac %.3f  launches/step %s clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['frac'], d['gpu_launches_per_step'], d['clocks']))
        print('train', d.get('train'))
        print('cpu', d.get('cpu_baseline'))
    else: print(l.rstrip())
"
tail -n 5 gpurun_out/bench.err
