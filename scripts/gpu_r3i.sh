#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "stay_inside_their_buffers" > gpurun_out/t_r3i.log 2>&1; echo "tests exit $?"; tail -n 25 gpurun_out/t_r3i.log | cut -c1-250
for e in "COORDX_BWD_PAIR=1" "COORDX_BWD_DFEAT_PAIR=0" "COORDX_BWD_FUSED=1" "COORDX_BWD_DZ_REUSE=0"; do
  env $e timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "stay_inside_their_buffers" > gpurun_out/t_r3i_b.log 2>&1; echo "$e exit $?"; tail -n 1 gpurun_out/t_r3i_b.log | cut -c1-200
done
