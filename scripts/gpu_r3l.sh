#!/bin/bash
# 3xTF32 GEMM, 128 x 256 tile as a CTA pair: correctness (whole GPU suite), then A/B on the video net
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_r3l.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/t_r3l.log | cut -c1-250
for v in 1 0 1 0; do
  echo "=== COORDX_G3_PAIR=$v fwd: $(COORDX_G3_PAIR=$v timeout 300 python scripts/kernel_breakdown.py cfg3_video_st_sum_h1024 fwd 2>&1 | grep -A1 'us of kernels' | cut -c1-110 | tr '\n' '|')"
  echo "=== COORDX_G3_PAIR=$v train: $(COORDX_G3_PAIR=$v timeout 300 python scripts/kernel_breakdown.py cfg3_video_st_sum_h1024 train 2>&1 | grep -A1 'us of kernels' | cut -c1-110 | tr '\n' '|')"
done
