#!/bin/bash
# end-metric parity: full-length runs (tests/convergence.py), then the reduced-step tests
mkdir -p gpurun_out
timeout 1500 python tests/convergence.py --log-every 2500 "$@" > gpurun_out/convergence.jsonl 2> gpurun_out/convergence.err; echo "converge exit $?"
cut -c1-1200 gpurun_out/convergence.jsonl; grep -v Warn gpurun_out/convergence.err | tail -n 24
