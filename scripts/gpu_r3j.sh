#!/bin/bash
# weight-gradient GEMMs on the tensor cores (dw_tensor: padded transposes + split-K 3xTF32 with an atomic epilogue)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_r3j.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/t_r3j.log | cut -c1-250
for cfg in cfg2_img2048_prod cfg3_video_st_sum_h1024; do
  for v in tc ffma; do
    if [ "$v" = "tc" ]; then unset COORDX_FP32_GEMM; else export COORDX_FP32_GEMM=ffma; fi
    echo "=== $cfg $v"; timeout 300 python scripts/kernel_breakdown.py $cfg train 2>&1 | grep -A8 "us of kernels" | cut -c1-120
  done
done
unset COORDX_FP32_GEMM
timeout 900 python scripts/bench_configs.py > gpurun_out/configs_r3j.jsonl 2> gpurun_out/configs_r3j.err; python - <<'P'
import json
for l in open('gpurun_out/configs_r3j.jsonl'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'fwd_ms', round(d.get('fwd_ms',0),4), 'train_graph_ms', d.get('train_graph_ms') and round(d['train_graph_ms'],4), 'train_ms', d.get('train_ms') and round(d['train_ms'],4))
P
