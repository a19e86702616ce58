#!/bin/bash
# Developer helper (run under gpurun): model parity tests + one 4096-tile bench with the per-kernel-family profile.
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --tiles 4096 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
  python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiles/s %.0f'%d['value']); [print('  %-16s %8.2f ms  %7.1f TF  %7.0f GB/s  x%d'%(k,v['ms'],v['tflops'],v['gbs'],v['launches'])) for k,v in d['kernels'].items() if v['ms']>0.5]"
