#!/bin/bash
# launch list of the tcgen05 trunk kernels of the training step (after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
$CMD > gpurun_out/plain_t.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpc__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:trunk_bwd -c 9 --csv --log-file gpurun_out/launches_bwd.csv $CMD > gpurun_out/ncu_t.log 2>&1
echo "launch list exit $?"
