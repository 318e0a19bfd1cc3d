#!/bin/bash
# one full ncu capture of the 128x64-tile 3xTF32 kernel (branch layer of the forward bench), after the same command ran plain
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-train --no-cpu"
$CMD > gpurun_out/plain_g3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 8 -c 1 -f -o gpurun_out/prof_g3_64 $CMD > gpurun_out/ncu_g3_64.log 2>&1
echo "full capture exit $?"; tail -n 2 gpurun_out/ncu_g3_64.log
