#!/bin/bash
# round 2, first GPU call: new backward kernels (dw2 + dfeat): parity tests, A/B timing against the v1 kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training" > gpurun_out/t_bwd.log 2>&1; echo "bwd tests exit $?"; tail -n 15 gpurun_out/t_bwd.log
timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train > gpurun_out/kb_v2.txt 2>&1; head -12 gpurun_out/kb_v2.txt
COORDX_BWD_V1=1 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train > gpurun_out/kb_v1.txt 2>&1; head -8 gpurun_out/kb_v1.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"; tail -n 15 gpurun_out/t_all.log
