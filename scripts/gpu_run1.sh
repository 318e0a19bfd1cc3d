#!/bin/bash
# first GPU pass: fp32 parity, then the tcgen05 path under a short timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== fp32 tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "fp32 or two_stage or errors" > gpurun_out/t_fp32.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/t_fp32.log
echo "=== tc tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "tensor_core or tc_" > gpurun_out/t_tc.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/t_tc.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/smoke.log
