#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 30 gpurun_out/t_all.log | cut -c1-250
