#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
$CMD > gpurun_out/plain_bwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${1:-trunk_bwd_dh} -s 2 -c 1 -f -o gpurun_out/prof_bwd $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "full capture exit $?"
