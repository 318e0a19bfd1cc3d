#!/bin/bash
# round 2: feature-gradient pass E2 variants, A/B (kernel time from torch.profiler)
mkdir -p gpurun_out
for v in default nodbl noset old2f default nodbl noset old2f; do
  if [ "$v" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$v.so; fi
  echo "=== lib=$v: $(timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dfeat' | cut -c1-70)"
done
