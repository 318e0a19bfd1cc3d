#!/bin/bash
# round 2: dW pass with the next sub-tile's TMEM load prefetched (A/B), then the ncu evidence of the training step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_backward_matches or tensor_core_training or falls_back" > gpurun_out/t_bwd.log 2>&1; echo "bwd tests exit $?"; tail -n 3 gpurun_out/t_bwd.log
for tag in default nopf default nopf; do
  if [ "$tag" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$tag.so; fi
  echo "=== $tag"
  timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A2 "us of kernels"
done
unset COORDX_B200_LIB
echo "=== default, COORDX_BWD_DZ_REUSE=0 (no dZ image)"
COORDX_BWD_DZ_REUSE=0 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -A3 "us of kernels"
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 2 gpurun_out/dw2_trace.txt
bash scripts/gpu_ll_bwd.sh
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
ncu --set full --clock-control none --import-source on -k regex:trunk_bwd_dw2 -s 2 -c 1 -f -o gpurun_out/prof_dw2 $CMD > gpurun_out/ncu_dw2.log 2>&1; echo "dw2 capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:trunk_bwd_dfeat -s 2 -c 1 -f -o gpurun_out/prof_dfeat $CMD > gpurun_out/ncu_dfeat.log 2>&1; echo "dfeat capture exit $?"
ls -la gpurun_out/*.ncu-rep
