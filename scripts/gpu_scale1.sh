#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_sharded.py 2>gpurun_out/sharded_n1.err > gpurun_out/sharded_n1.jsonl; echo "sharded N=1 exit $?"
cut -c1-500 gpurun_out/sharded_n1.jsonl; tail -n 3 gpurun_out/sharded_n1.err
