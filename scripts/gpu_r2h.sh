#!/bin/bash
# round 2: dW pass with two E1 groups on alternate sub-tiles (A/B against all 16 warps per sub-tile)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training or falls_back or tc_loss" > gpurun_out/t_bwd.log 2>&1; echo "bwd tests exit $?"; tail -n 3 gpurun_out/t_bwd.log
for tag in default noalt default noalt; do
  if [ "$tag" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$tag.so; fi
  echo "=== $tag"
  timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A2 "us of kernels"
done
unset COORDX_B200_LIB
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 14 gpurun_out/dw2_trace.txt
