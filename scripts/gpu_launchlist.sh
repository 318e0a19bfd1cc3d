#!/bin/bash
# launch list of the forward bench (after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-train --no-cpu"
$CMD > gpurun_out/plain_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_fwd.csv $CMD > gpurun_out/ncu_ll.log 2>&1
echo "launch list exit $?"
