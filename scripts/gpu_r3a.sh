#!/bin/bash
# E1 loop with the sub-tile half as a compile-time constant (unr) vs current; plus merged preps; new tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "kernel_forms or tc_backward_matches or tensor_core_training or falls_back or captured_train" > gpurun_out/t_r3a.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_r3a.log | cut -c1-250
export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_unr.so
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "kernel_forms or tc_backward_matches or tensor_core_training" > gpurun_out/t_r3a_unr.log 2>&1; echo "unr tests exit $?"; tail -n 2 gpurun_out/t_r3a_unr.log | cut -c1-250
for v in default unr default unr; do
  if [ "$v" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$v.so; fi
  echo "=== lib=$v: $(timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dw2' | cut -c1-70)"
done
unset COORDX_B200_LIB
