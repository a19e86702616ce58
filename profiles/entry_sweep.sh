#!/bin/bash
# Developer helper (run under gpurun): throughput vs entry-flow sub-batch size (L2 residency of the 147^2..37^2 maps).
for eb in 4 8 16 32 0; do
  echo "=== entry_batch $eb"
  BQ_ENTRY_BATCH=$eb timeout 300 python bench.py --tiles 2048 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiles/s %.0f'%d['value']); [print('  %-16s %8.2f ms  %7.1f TF  %7.0f GB/s  x%d'%(k,v['ms'],v['tflops'],v['gbs'],v['launches'])) for k,v in d['kernels'].items() if v['ms']>1.5]"
done
