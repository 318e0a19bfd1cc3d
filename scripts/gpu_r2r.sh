#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dfeat_trace.py > gpurun_out/dfeat_trace.txt 2>&1; tail -n 30 gpurun_out/dfeat_trace.txt | cut -c1-220
timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dfeat' | cut -c1-70
