#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
$CMD > gpurun_out/plain_bwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trunk_bwd_tc -s 2 -c 2 -f -o gpurun_out/prof_bwd $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "full capture exit $?"
tail -n 2 gpurun_out/ncu_bwd.log
