#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 5 --warmup 3 --train-steps 12 2>gpurun_out/bench_n2.err > gpurun_out/bench_n2.json; echo "bench N=2 exit $?"
grep "uneven\|Error\|error" gpurun_out/bench_n2.err | head
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=%d value %.3e ms %.4f e2e %.3e train %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['train']))
"
