#!/bin/bash
# round 2: fused backward kernel: relaxed vs release hand-over, producer-only timing (negative chunk), role split
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_backward" > gpurun_out/t_fused.log 2>&1; echo "fused tests exit $?"; tail -n 2 gpurun_out/t_fused.log
run() {  # pairs chunk groups
  COORDX_BWD_FUSED_PAIRS=$1 COORDX_BWD_FUSED_CHUNK=$2 COORDX_BWD_FUSED_GROUPS=$3 timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep "trunk_bwd_fused" | cut -c1-60
}
for lib in default rel; do
  if [ "$lib" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$lib.so; fi
  echo "##### lib=$lib"
  for cfg in "56 4 18" "56 -4 18" "73 -4 1" "73 -512 1" "58 4 16" "60 4 14" "62 4 12" "60 8 14"; do
    set -- $cfg
    echo "=== pairs=$1 chunk=$2 groups=$3: $(run $1 $2 $3)"
  done
done
