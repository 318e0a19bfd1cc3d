#!/bin/bash
# 3xTF32 GEMM: accuracy test, the fp32-path parity tests that now route through it, and the video configuration's timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu -k "tf32x3 or separable or fp32" -s > gpurun_out/t_tf32.log 2>&1; echo "tests exit $?"; grep -E "rel err|passed|failed|Error" gpurun_out/t_tf32.log | tail -n 8
timeout 600 python scripts/bench_configs.py --only cfg3_video_st_sum_h1024 --steps 5 2>gpurun_out/cfg3.err | tee gpurun_out/cfg3.jsonl | cut -c1-600
COORDX_FP32_GEMM=ffma timeout 600 python scripts/bench_configs.py --only cfg3_video_st_sum_h1024 --steps 3 2>>gpurun_out/cfg3.err | tee -a gpurun_out/cfg3.jsonl | cut -c1-300
timeout 300 python scripts/kernel_breakdown.py cfg3_video_st_sum_h1024 fwd 2>&1 | tail -n 25
