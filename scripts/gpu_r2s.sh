#!/bin/bash
# round 2: feature-gradient pass as CTA pairs (trunk_bwd_dfeat_pair_kernel): correctness, A/B, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training or falls_back or fused_backward or captured_train or large_grid or scattered" > gpurun_out/t_dfp.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_dfp.log | cut -c1-300
for v in 1 0 1 0; do
  echo "=== COORDX_BWD_DFEAT_PAIR=$v: $(COORDX_BWD_DFEAT_PAIR=$v timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dfeat' | cut -c1-80)"
done
timeout 300 python scripts/dfeat_trace.py > gpurun_out/dfeat_pair_trace.txt 2>&1; tail -n 24 gpurun_out/dfeat_pair_trace.txt | cut -c1-220
