#!/bin/bash
# round 2: dW pass as CTA pairs (trunk_bwd_dw3_kernel, cta_group::2): correctness, then A/B against the single-CTA form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training or falls_back or fused_backward or captured_train" > gpurun_out/t_pair.log 2>&1; echo "pair tests exit $?"; tail -n 5 gpurun_out/t_pair.log | cut -c1-300
for pair in 1 0 1 0; do
  echo "=== COORDX_BWD_PAIR=$pair"
  COORDX_BWD_PAIR=$pair timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels" | cut -c1-110
done
COORDX_BWD_PAIR=1 timeout 300 python scripts/dw2_trace.py > gpurun_out/dw3_trace.txt 2>&1; tail -n 42 gpurun_out/dw3_trace.txt | cut -c1-200
