#!/bin/bash
# round 2: feature-gradient pass with double-buffered TMEM loads + packed arithmetic in E2 (E2 at 112 registers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training or falls_back or fused_backward or captured_train" > gpurun_out/t_df.log 2>&1; echo "tests exit $?"; tail -n 2 gpurun_out/t_df.log | cut -c1-300
for i in 1 2; do
  timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels" | cut -c1-110
done
timeout 300 python scripts/kernel_breakdown.py cfg4_occ256_prod train 2>&1 | grep -A4 "us of kernels" | cut -c1-110
