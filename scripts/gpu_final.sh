#!/bin/bash
# round-end evidence: full GPU tests, smoke, default bench, reference arm, launch lists (forward + training)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench_n1.err > gpurun_out/bench_n1.json; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_ll.log 2>&1
echo "launch list exit $?"
