#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 6 gpurun_out/t_all.log
timeout 300 python scripts/train_timeline.py cfg2_img2048_prod 2>/dev/null > gpurun_out/train_timeline_cfg2.txt; echo "exit $?"
head -3 gpurun_out/train_timeline_cfg2.txt; grep -n "trunk_bwd_din" -A40 gpurun_out/train_timeline_cfg2.txt | cut -c1-120
timeout 300 python scripts/train_timeline.py cfg1_img256_sum 2>/dev/null > gpurun_out/train_timeline_cfg1.txt; echo "exit $?"
cut -c1-120 gpurun_out/train_timeline_cfg1.txt
