#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py --steps 10 2>gpurun_out/configs.err | tee gpurun_out/configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], 'fwd_ms', d['fwd_ms'], 'train_ms', d.get('train_ms'), 'graph_ms', d.get('train_graph_ms'))
"
timeout 300 python scripts/bench_unsplit.py 2>/dev/null > gpurun_out/unsplit.json; cut -c1-330 gpurun_out/unsplit.json
