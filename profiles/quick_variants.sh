#!/bin/bash
# Developer helper (run under gpurun): parity tests of the model path, then a short bench per kernel variant.
set -u
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -4
for v in "default"; do
  echo "=== $v"
  if [ "$v" = "default" ]; then envs=""; else envs="$v"; fi
  env $envs timeout 300 python bench.py --tiles 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiles/s %.0f'%d['value']); [print('  %-16s %8.2f ms  %7.1f TF  %7.0f GB/s  x%d'%(k,v['ms'],v['tflops'],v['gbs'],v['launches'])) for k,v in d['kernels'].items()]"
done
