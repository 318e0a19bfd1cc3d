#!/bin/bash
# round 2: forward epilogue with a share of the sines on the FMA pipe (packed fp32x2 polynomial): A/B of the share
mkdir -p gpurun_out
for tag in fp4 fp5 fp6 fp8 fp4 fp5 fp6 fp8; do
  if [ "$tag" = "default" ]; then unset COORDX_B200_LIB; else export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_$tag.so; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-train > gpurun_out/b_$tag.json 2>/dev/null
  python - "$tag" <<'P'
import json, sys
d=json.loads(open(f'gpurun_out/b_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.3f G/s' % (d['value']/1e9), 'kernel %.4f ms' % d['roofline']['kernel_ms'], 'achieved %.0f TF' % d['roofline']['achieved'], 'e2e %.3f' % (d['e2e']['value']/1e9), d['clocks']['sm_mhz'])
P
done
export COORDX_B200_LIB=$PWD/coordx_b200/_build/libcoordx_fp6.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_forward or tensor_core or concurrent" > gpurun_out/t_fp6.log 2>&1; echo "fp6 tests exit $?"; tail -n 3 gpurun_out/t_fp6.log
