#!/bin/bash
# A/B: weight-gradient / reduction GEMMs on the tensor cores (default) vs FFMA (COORDX_FP32_GEMM=ffma), bench.py training key
mkdir -p gpurun_out
for v in tc ffma tc ffma; do
  if [ "$v" = "tc" ]; then unset COORDX_FP32_GEMM; else export COORDX_FP32_GEMM=ffma; fi
  timeout 600 python bench.py --no-cpu > gpurun_out/b_x.json 2>/dev/null
  python - "$v" <<'P'
import json,sys
d=json.loads(open('gpurun_out/b_x.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'train ms', round(d['train']['ms_per_iter'],4), 'frac', round(d['train']['frac_of_sustained_peak'],4), 'fwd G/s', round(d['value']/1e9,3))
P
done
