for mb in 503 500 503 500; do
  timeout 200 python bench.py --steps 3 --warmup 3 --max-batch $mb --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('max_batch $mb tiles/s %.0f' % d['value'], ' '.join('%s=%.1f' % (n, k[n]['ms']) for n in ('sepconv_mid','gemm_pointwise','sepconv_fused') if n in k))"
done
