#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dw2_trace.py > gpurun_out/dw2_trace.txt 2>&1; tail -n 50 gpurun_out/dw2_trace.txt
