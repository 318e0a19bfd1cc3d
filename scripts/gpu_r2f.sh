#!/bin/bash
# round 2: forward constant copies (4 TU copies) check + dW-pass A/B (polynomial share, E1 warp skew) + fixed micro-benchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "concurrent or fp16_relu or forward_matches or tc_forward or tensor_core or empty" > gpurun_out/t_fwd.log 2>&1; echo "fwd tests exit $?"; tail -n 3 gpurun_out/t_fwd.log
timeout 120 scripts/micro/micro_e1 > gpurun_out/micro_e1.txt 2>&1; cat gpurun_out/micro_e1.txt
for tag in default p2 p3 k300 k600 k900; do
  if [ "$tag" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$tag.so; fi
  echo "=== $tag"
  timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels"
done
export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_k600.so
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace_k600.txt 2>&1; tail -n 12 gpurun_out/dw2_trace_k600.txt
unset COORDX_B200_LIB
timeout 600 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2f.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel_ms'], d['train']['value'], d['train']['ms_per_iter'])
P
