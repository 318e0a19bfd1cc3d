#!/bin/bash
# round 2: new defaults (forward: 4 of 16 column pairs on the FMA-pipe polynomial; dW pass: 2 of 8) - full suite, timeline, bench, ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"; tail -n 3 gpurun_out/t_all.log
timeout 300 python scripts/fwd_trace.py > gpurun_out/fwd_trace.txt 2>&1; tail -n 12 gpurun_out/fwd_trace.txt
timeout 600 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench exit $?"; cat gpurun_out/bench_r2j.json | cut -c1-600
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --train-steps 2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_ll.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:trunk_fwd_tc -s 4 -c 1 -f -o gpurun_out/prof_fwd $CMD > gpurun_out/ncu_fwd.log 2>&1; echo "fwd capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:trunk_bwd_dw2 -s 2 -c 1 -f -o gpurun_out/prof_dw2 $CMD > gpurun_out/ncu_dw2.log 2>&1; echo "dw2 capture exit $?"
