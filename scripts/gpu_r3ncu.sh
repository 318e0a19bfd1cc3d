#!/bin/bash
# final-form captures: dW pass after the E1 trims, and the CTA-pair 3xTF32 GEMM of the video net (summaries come back, the .ncu-rep stay)
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
$CMD > gpurun_out/plain.log 2>&1; echo "plain exit $?"
ncu --set full --clock-control none --import-source on -k regex:trunk_bwd_dw2 -s 2 -c 1 -f -o /tmp/ncu/prof_dw2 $CMD > gpurun_out/ncu_dw2.log 2>&1; echo "dw2 capture exit $?"
python scripts/summarize_ncu.py full /tmp/ncu/prof_dw2.ncu-rep > gpurun_out/r02s_dw2_full.txt 2>&1
CMD2="python scripts/kernel_breakdown.py cfg3_video_st_sum_h1024 fwd"
$CMD2 > gpurun_out/plain2.log 2>&1; echo "plain2 exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_pair -s 30 -c 1 -f -o /tmp/ncu/prof_g3p $CMD2 > gpurun_out/ncu_g3p.log 2>&1; echo "g3 pair capture exit $?"
python scripts/summarize_ncu.py full /tmp/ncu/prof_g3p.ncu-rep > gpurun_out/r02s_g3pair_full.txt 2>&1
tail -n 30 gpurun_out/r02s_g3pair_full.txt | cut -c1-150
