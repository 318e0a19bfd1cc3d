#!/bin/bash
# A/B: dfeat pair issue-ahead limit (ah2, ah1), forward epilogue batch split (eb2, eb4) after the trace removal
mkdir -p gpurun_out
for v in default ah2 ah1 default ah2 ah1; do
  if [ "$v" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$v.so; fi
  echo "=== lib=$v: $(timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dfeat' | cut -c1-70)"
done
for v in default eb2 eb4 default eb2 eb4; do
  if [ "$v" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$v.so; fi
  timeout 300 python bench.py --no-cpu --no-train --steps 30 --warmup 5 > gpurun_out/b_x.json 2>/dev/null
  python - "$v" <<'P'
import json,sys
d=json.loads(open('gpurun_out/b_x.json').read().strip().splitlines()[-1])
print('lib', sys.argv[1], 'fwd G/s', round(d['value']/1e9,3), 'kernel ms', round(d['roofline']['kernel_ms'],4), 'TF', round(d['roofline']['achieved'],1), 'sustained ms', round(d['roofline']['sustained_ms_per_step'],4))
P
done
