#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dataio.py -q -x > gpurun_out/t_dataio.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/t_dataio.log | cut -c1-250
timeout 600 python scripts/bench_video_batch.py > gpurun_out/video_batch.json 2> gpurun_out/video_batch.err; echo "bench exit $?"; cat gpurun_out/video_batch.json; tail -n 3 gpurun_out/video_batch.err
