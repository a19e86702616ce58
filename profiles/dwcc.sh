#!/bin/bash
for v in 64 32; do
  echo "=== BQ_DW_CC=$v"
  BQ_DW_CC=$v timeout 300 python bench.py --tiles 4096 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiles/s %.0f'%d['value']); [print('  %-16s %8.2f ms  %7.1f TF  %7.0f GB/s  x%d'%(k,v['ms'],v['tflops'],v['gbs'],v['launches'])) for k,v in d['kernels'].items() if v['ms']>2]"
done
