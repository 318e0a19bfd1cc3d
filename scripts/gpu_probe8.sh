#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_all.log | cut -c1-300
for c in cfg2_img2048_prod cfg1_img256_sum; do
timeout 300 python scripts/train_timeline.py $c 2>/dev/null > gpurun_out/train_timeline_$c.txt; echo "exit $?"
head -1 gpurun_out/train_timeline_$c.txt; done
head -14 gpurun_out/train_timeline_cfg2_img2048_prod.txt | cut -c1-120
grep -n "trunk_bwd_din" -A34 gpurun_out/train_timeline_cfg2_img2048_prod.txt | cut -c1-120
timeout 900 python scripts/bench_configs.py --steps 10 2>gpurun_out/configs.err | tee gpurun_out/configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], 'fwd_ms', d['fwd_ms'], 'train_ms', d.get('train_ms'), 'graph_ms', d.get('train_graph_ms'))
"
