#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_unsplit.py 2>gpurun_out/unsplit.err | tee gpurun_out/unsplit.json | cut -c1-900; tail -n 3 gpurun_out/unsplit.err
