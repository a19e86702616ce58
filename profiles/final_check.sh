#!/bin/bash
# Developer helper (run under gpurun): whole GPU suite, then the per-kernel profile at micro-batch 256 and 512.
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for b in 256 512; do
  echo "=== max-batch $b"
  timeout 300 python bench.py --tiles 6144 --steps 2 --warmup 3 --max-batch $b --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiles/s %.0f'%d['value']); [print('  %-16s %8.2f ms  %7.1f TF  %7.0f GB/s  x%d'%(k,v['ms'],v['tflops'],v['gbs'],v['launches'])) for k,v in d['kernels'].items() if v['ms']>0.5]"
done
