#!/bin/bash
# fp32-precision trunk: dX of the first trunk layer on the tensor cores, one dW pass per 32768-point chunk, vectorised fusion kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_r3o.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r3o.log | cut -c1-250
COORDX_PRECISION=fp32 timeout 600 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A10 "us of kernels" | cut -c1-140
timeout 300 python scripts/kernel_breakdown.py cfg3_video_st_sum_h1024 train 2>&1 | grep -A4 "us of kernels" | cut -c1-140
