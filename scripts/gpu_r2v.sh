#!/bin/bash
# round 2: loss fused into the trunk's last layer: parity, then bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "loss_fused or captured_train or fused_loss or runtime_graph" > gpurun_out/t_loss.log 2>&1; echo "tests exit $?"; tail -n 12 gpurun_out/t_loss.log | cut -c1-250
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r02k.json').read().strip().splitlines()[-1])
print('fwd', d['value']/1e9, 'kernel_ms', d['roofline']['kernel_ms'], 'achieved', d['roofline']['achieved'], 'e2e', d['e2e']['value']/1e9, 'train', d['train']['value'], d['train']['ms_per_iter'], d['train']['frac_of_sustained_peak'], d['clocks'])
P
timeout 300 python scripts/train_timeline.py cfg2_img2048_prod > gpurun_out/train_timeline_r02k.txt 2>&1; grep -v "^/opt\|_warn" gpurun_out/train_timeline_r02k.txt | sed -n 1,30p | cut -c1-120
