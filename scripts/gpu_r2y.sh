#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "loss_fused or captured_train" > gpurun_out/t_loss.log 2>&1; echo "tests exit $?"; tail -n 2 gpurun_out/t_loss.log | cut -c1-250
for v in 1 0 1 0 1 0; do
  COORDX_FUSED_LOSS=$v timeout 600 python bench.py --no-cpu > gpurun_out/b_x.json 2>/dev/null
  python - "$v" <<'P'
import json,sys
d=json.loads(open('gpurun_out/b_x.json').read().strip().splitlines()[-1])
print('COORDX_FUSED_LOSS', sys.argv[1], 'train ms', round(d['train']['ms_per_iter'],4), 'frac', round(d['train']['frac_of_sustained_peak'],4), 'fwd G/s', round(d['value']/1e9,3))
P
done
