#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 12 gpurun_out/t_all.log | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 scripts/bench_sharded.py 2>gpurun_out/sharded_n2.err > gpurun_out/sharded_n2.jsonl; echo "sharded N=2 exit $?"
cut -c1-900 gpurun_out/sharded_n2.jsonl; grep -i "error\|Traceback" -A5 gpurun_out/sharded_n2.err | head -20
