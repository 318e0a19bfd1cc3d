#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core_training" > gpurun_out/t_bwd.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/t_bwd.log
