#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
$CMD > gpurun_out/plain_t.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_t.log 2>&1
echo "launch list exit $?"
