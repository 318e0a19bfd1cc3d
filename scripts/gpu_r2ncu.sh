#!/bin/bash
# round 2, final kernels: launch list of the default bench command + ncu --set full captures of the three trunk kernels.
# The .ncu-rep files stay on the box (3 x 30 MB > gpurun's 64 MiB return limit): summaries + raw metric CSVs come back.
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --train-steps 2"
$CMD > gpurun_out/plain.log 2>&1; echo "plain exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02m.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_ll.log 2>&1; echo "launch list exit $?"
for k in "fwd trunk_fwd_tc_kernel 4" "dw2 trunk_bwd_dw2 2" "dfeat trunk_bwd_dfeat_pair 2"; do
  set -- $k
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/ncu/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1; echo "$1 capture exit $?"
  python scripts/summarize_ncu.py full /tmp/ncu/prof_$1.ncu-rep > gpurun_out/r02m_$1_full.txt 2>&1
  ncu -i /tmp/ncu/prof_$1.ncu-rep --page raw --csv > gpurun_out/r02m_$1_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/prof_$1.ncu-rep --page details --csv > gpurun_out/r02m_$1_details.csv 2>/dev/null
done
ls -la gpurun_out | head -30
