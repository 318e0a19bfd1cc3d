#!/bin/bash
for NM in 32 30 28; do
  LIB=coordx_b200/_build/libcoordx_nm$NM.so
  [ $NM = 32 ] && LIB=coordx_b200/libcoordx_b200.so
  echo "== NM=$NM"
  COORDX_B200_LIB=$PWD/$LIB timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('ms/step %.3f kernel_ms %.4f achieved %.1f TF frac %.3f sustained %.3f' % (d['ms_per_step'], r['kernel_ms'], r['achieved'], r['frac'], r['sustained_ms_per_step']))
"
done
