#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "slabs_through_the_c_abi" > gpurun_out/t_r3e.log 2>&1; echo "tests exit $?"; tail -n 25 gpurun_out/t_r3e.log | cut -c1-250
