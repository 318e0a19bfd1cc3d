#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu --deselect tests/test_z_convergence_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -n 20 gpurun_out/t_all.log | cut -c1-250
timeout 600 python bench.py --no-cpu 2>gpurun_out/bench_fwd.err > gpurun_out/bench_fwd.json; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_fwd.json').read().strip().splitlines()[-1])
print('value %.3e ms %.4f e2e %.3e frac %.3f train %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['train']['value']))
"
timeout 300 python scripts/bench_unsplit.py 2>/dev/null | cut -c1-400
timeout 300 python scripts/train_timeline.py cfg2_img2048_prod 2>/dev/null > gpurun_out/train_timeline_cfg2_img2048_prod.txt; head -12 gpurun_out/train_timeline_cfg2_img2048_prod.txt | cut -c1-110
