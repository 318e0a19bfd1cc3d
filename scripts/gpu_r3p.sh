#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_r3p.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r3p.log | cut -c1-250
COORDX_PRECISION=fp32 timeout 600 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A6 "us of kernels" | cut -c1-140
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
