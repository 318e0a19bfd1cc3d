#!/bin/bash
# round 2: dW pass as CTA pairs: forwarded vs direct remote arrives vs single-CTA form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training or falls_back or fused_backward or captured_train" > gpurun_out/t_pair.log 2>&1; echo "pair tests exit $?"; tail -n 2 gpurun_out/t_pair.log | cut -c1-300
for v in "default 1" "nofwd 1" "default 0" "default 1" "nofwd 1" "default 0"; do
  set -- $v
  if [ "$1" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$1.so; fi
  echo "=== lib=$1 COORDX_BWD_PAIR=$2: $(COORDX_BWD_PAIR=$2 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dw' | cut -c1-70)"
done
unset COORDX_B200_LIB
COORDX_BWD_PAIR=1 timeout 300 python scripts/dw2_trace.py > gpurun_out/dw3_trace.txt 2>&1; tail -n 12 gpurun_out/dw3_trace.txt | cut -c1-200
for dbg in 0 1 2 3 4 5 6 7; do
  echo "=== dfeat COORDX_BWD_DBG=$dbg (1 no E2 math, 2 no MMA, 4 no loads): $(COORDX_BWD_PAIR=0 COORDX_BWD_DBG=$dbg timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep 'trunk_bwd_dfeat' | cut -c1-70)"
done
