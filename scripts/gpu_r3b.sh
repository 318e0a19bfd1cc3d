#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "kernel_forms or tc_backward_matches or tensor_core_training or falls_back or captured_train or fused_backward or loss_fused or smoke" > gpurun_out/t_r3b.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r3b.log | cut -c1-250
for v in default unr2 default unr2; do
  if [ "$v" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$v.so; fi
  echo "=== lib=$v: $(timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_d' | cut -c1-70 | tr '\n' '|')"
done
unset COORDX_B200_LIB
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r3b.json 2> gpurun_out/bench_r3b.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r3b.json').read().strip().splitlines()[-1])
print('fwd', d['value']/1e9, 'kernel_ms', d['roofline']['kernel_ms'], 'achieved', d['roofline']['achieved'], 'e2e', d['e2e']['value']/1e9, 'train', d['train']['value'], d['train']['ms_per_iter'], d['train']['frac_of_sustained_peak'])
P
