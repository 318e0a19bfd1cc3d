#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/kernel_breakdown.py cfg2_img2048_prod train 2>&1 | grep -v Warn | tail -n 28 | tee gpurun_out/kb_cfg2_train.txt
CMD="python scripts/bench_configs.py --only cfg3_video_st_sum_h1024 --steps 2"
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 6 -c 1 -f -o gpurun_out/prof_g3 $CMD > gpurun_out/ncu_g3.log 2>&1
echo "full capture exit $?"; tail -n 2 gpurun_out/ncu_g3.log
