#!/bin/bash
# round 2: validate HEAD on the GPU: all gpu tests, training-step kernel breakdown, dW-pass trace, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"; tail -n 6 gpurun_out/t_all.log
timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train > gpurun_out/kb_head.txt 2>&1; head -8 gpurun_out/kb_head.txt
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 4 gpurun_out/dw2_trace.txt
timeout 600 python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "bench exit $?"; cat gpurun_out/bench_head.json
