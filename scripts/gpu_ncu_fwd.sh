#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-train --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trunk_fwd_tc -s 4 -c 2 -o gpurun_out/prof_fwd $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?"
tail -n 3 gpurun_out/ncu2.log
