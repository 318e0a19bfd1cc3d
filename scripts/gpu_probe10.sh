#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 25 gpurun_out/t_all.log | cut -c1-250
timeout 900 python scripts/bench_configs.py --steps 10 2>gpurun_out/configs.err | tee gpurun_out/configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], 'fwd_ms', d['fwd_ms'], 'train_ms', d.get('train_ms'), 'graph_ms', d.get('train_graph_ms'))
"
timeout 300 python tests/convergence.py --only img512_prod_fp32 --steps 300 2>/dev/null | cut -c1-500
