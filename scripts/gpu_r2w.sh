#!/bin/bash
# A/B: loss fused into the trunk vs separate loss kernel (captured training step, bench.py train key)
mkdir -p gpurun_out
for v in 1 0 1 0; do
  COORDX_FUSED_LOSS=$v timeout 600 python bench.py --no-cpu > gpurun_out/b_loss$v.json 2>/dev/null
  python - $v <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/b_loss{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('COORDX_FUSED_LOSS', sys.argv[1], 'train ms', d['train']['ms_per_iter'], 'frac', d['train']['frac_of_sustained_peak'], 'fwd kernel ms', d['roofline']['kernel_ms'])
P
done
