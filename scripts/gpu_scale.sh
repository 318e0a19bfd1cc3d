#!/bin/bash
# N-GPU run: $1 = N.  bench.py (weak scaling, the driver's contract) + the sharded 4K / 512^3 workloads (strong scaling)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json; echo "bench N=$N exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=%d value %.3e ms %.4f e2e %.3e train %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['train'] and d['train']['value']))
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29502 scripts/bench_sharded.py 2>gpurun_out/sharded_n$N.err > gpurun_out/sharded_n$N.jsonl; echo "sharded N=$N exit $?"
cut -c1-500 gpurun_out/sharded_n$N.jsonl; tail -n 3 gpurun_out/sharded_n$N.err
