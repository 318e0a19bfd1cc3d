#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
