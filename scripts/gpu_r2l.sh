#!/bin/bash
# round 2: fused backward kernel (dW CTAs + feature-gradient CTAs in one launch): correctness, then the role-split sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_backward or tc_backward_matches or tensor_core_training or falls_back" > gpurun_out/t_fused.log 2>&1; echo "fused tests exit $?"; tail -n 5 gpurun_out/t_fused.log
echo "=== two passes (COORDX_BWD_FUSED=0)"
COORDX_BWD_FUSED=0 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels"
for cfg in "56 4 18" "52 4 22" "54 4 20" "58 4 16" "60 4 14" "56 2 18" "56 8 18" "56 4 9" "56 4 36" "55 4 19"; do
  set -- $cfg
  echo "=== fused pairs=$1 chunk=$2 groups=$3"
  COORDX_BWD_FUSED_PAIRS=$1 COORDX_BWD_FUSED_CHUNK=$2 COORDX_BWD_FUSED_GROUPS=$3 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A1 "us of kernels"
done
timeout 300 python scripts/train_timeline.py cfg2_img2048_prod > gpurun_out/train_timeline_r02f.txt 2>&1; tail -n 70 gpurun_out/train_timeline_r02f.txt | cut -c1-150
