#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/train_timeline.py cfg2_img2048_prod 2>/dev/null > gpurun_out/train_timeline_cfg2.txt; echo "exit $?"
cat gpurun_out/train_timeline_cfg2.txt | cut -c1-150
