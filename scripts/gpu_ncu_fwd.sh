#!/bin/bash
# one full ncu capture of the forward trunk kernel (after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-train --no-cpu"
$CMD > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trunk_fwd_tc -s 4 -c 1 -f -o gpurun_out/prof_fwd $CMD > gpurun_out/ncu_fwd.log 2>&1
echo "full capture exit $?"
tail -n 3 gpurun_out/ncu_fwd.log
