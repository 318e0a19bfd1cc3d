#!/bin/bash
# fused-loss cost breakdown (timing-only variants: results of ld* are wrong)
mkdir -p gpurun_out
for v in "default 1" "default 0" "ld1 1" "ld2 1" "ld4 1" "ld7 1" "default 1" "default 0"; do
  set -- $v
  if [ "$1" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$1.so; fi
  COORDX_FUSED_LOSS=$2 timeout 600 python bench.py --no-cpu > gpurun_out/b_x.json 2>/dev/null
  python - "$1" "$2" <<'P'
import json,sys
d=json.loads(open('gpurun_out/b_x.json').read().strip().splitlines()[-1])
print('lib', sys.argv[1], 'COORDX_FUSED_LOSS', sys.argv[2], 'train ms', round(d['train']['ms_per_iter'],4), 'frac', round(d['train']['frac_of_sustained_peak'],4))
P
done
