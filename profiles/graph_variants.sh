#!/bin/bash
# Developer helper (run under gpurun): CUDA-graph replay of the backbone (default) vs plain stream launches.
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
for v in on off on off; do
  echo "=== BQ_GRAPH=$v"
  BQ_GRAPH=$v timeout 300 python bench.py --tiles 6144 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tiles/s %.0f  launches %d'%(d['value'], d['gpu_launches']))"
done
