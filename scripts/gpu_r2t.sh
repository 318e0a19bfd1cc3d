#!/bin/bash
mkdir -p gpurun_out
for v in default notrace default notrace; do
  if [ "$v" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$v.so; fi
  echo "=== lib=$v"; timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd' | cut -c1-80
done
unset COORDX_B200_LIB
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r02i.json').read().strip().splitlines()[-1])
print('fwd', d['value']/1e9, 'train', d['train']['value'], d['train']['ms_per_iter'], d['train']['frac_of_sustained_peak'])
P
