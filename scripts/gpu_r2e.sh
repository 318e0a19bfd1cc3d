#!/bin/bash
# round 2: robustness fixes (per-device attributes, constant slots, chunked fp32 trunk, dZ fallback, activations) + dW-pass A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"; tail -n 6 gpurun_out/t_all.log
timeout 120 scripts/micro/micro_e1 > gpurun_out/micro_e1.txt 2>&1; cat gpurun_out/micro_e1.txt
for tag in default e0 p2 p3 p5 s4 s6; do
  if [ "$tag" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$tag.so; fi
  echo "=== $tag"
  timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels"
done
unset COORDX_B200_LIB
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 3 gpurun_out/dw2_trace.txt
timeout 600 python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel_ms'], d['train']['value'], d['train']['ms_per_iter'])
P
