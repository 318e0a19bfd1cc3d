#!/bin/bash
# round 2: tracing compiled out of the product library + pair feature-gradient pass: full GPU tests, bench, timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r02j.json').read().strip().splitlines()[-1])
print('fwd', d['value']/1e9, 'kernel_ms', d['roofline']['kernel_ms'], 'achieved', d['roofline']['achieved'], 'e2e', d['e2e']['value']/1e9, 'train', d['train']['value'], d['train']['ms_per_iter'], d['train']['frac_of_sustained_peak'], d['clocks'])
P
timeout 300 python scripts/train_timeline.py cfg2_img2048_prod > gpurun_out/train_timeline_r02j.txt 2>&1; grep -c . gpurun_out/train_timeline_r02j.txt
timeout 300 python scripts/dfeat_trace.py > gpurun_out/dfeat_pair_trace.txt 2>&1; tail -n 4 gpurun_out/dfeat_pair_trace.txt | cut -c1-220
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 2 gpurun_out/dw2_trace.txt | cut -c1-220
