#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 5 gpurun_out/t_all.log | cut -c1-300
for c in cfg2_img2048_prod cfg1_img256_sum; do
timeout 300 python scripts/train_timeline.py $c 2>/dev/null > gpurun_out/train_timeline_$c.txt; echo "exit $?"
head -1 gpurun_out/train_timeline_$c.txt; done
grep -n "g3::\|transpose" gpurun_out/train_timeline_cfg2_img2048_prod.txt | cut -c1-120
grep -n "g3::" gpurun_out/train_timeline_cfg1_img256_sum.txt | cut -c1-120 | head
timeout 600 python scripts/bench_configs.py --only cfg3 --steps 5 2>/dev/null | cut -c1-400
