#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core_training" > gpurun_out/t_bwd.log 2>&1; echo "tests exit $?"; tail -n 25 gpurun_out/t_bwd.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --train-steps 10 2>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('fwd ms', d['ms_per_step'], 'train', d.get('train'))
"
tail -n 3 gpurun_out/bench.err
