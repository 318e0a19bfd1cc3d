#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels on small grids (smoke + the kernel-form test + the loss-fused test + dataio)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python __graft_entry__.py smoke > gpurun_out/san_smoke.log 2>&1; echo "memcheck smoke exit $?"; grep -c "Invalid\|Error" gpurun_out/san_smoke.log; tail -n 3 gpurun_out/san_smoke.log | cut -c1-200
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py tests/test_dataio.py -q -x -k "kernel_forms or gpu_batches_match or (loss_fused and cfg1)" > gpurun_out/san_tests.log 2>&1; echo "memcheck tests exit $?"; grep -c "Invalid\|out of bounds" gpurun_out/san_tests.log; tail -n 4 gpurun_out/san_tests.log | cut -c1-200
