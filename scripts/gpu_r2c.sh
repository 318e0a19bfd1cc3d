#!/bin/bash
# time the training step's kernels with several builds of the backward TU (COORDX_B200_LIB override)
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$tag.so; fi
  echo "=== $tag"
  timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels"
done
unset COORDX_B200_LIB
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 40 gpurun_out/dw2_trace.txt
