#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/t_all.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit $?"; tail -n 3 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
