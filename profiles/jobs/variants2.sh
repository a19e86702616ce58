cp biscuit_b200/libbiscuit_b200.so /tmp/lib_default.so
for f in profiles/variants/lib_g*.so; do
  v=$(basename $f .so)
  cp $f biscuit_b200/libbiscuit_b200.so
  timeout 200 python bench.py --tiles 4096 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> /tmp/err_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$v', 'tiles/s %.0f' % d['value'], ' '.join('%s=%.1f' % (n, k[n]['ms']) for n in ('sepconv_mid','maxpool_add','subsample','depthwise','gemm_pointwise') if n in k))"
  tail -2 /tmp/err_$v.log | cut -c1-300
done
cp /tmp/lib_default.so biscuit_b200/libbiscuit_b200.so
# second pass in the opposite order (box drift)
for f in $(ls -r profiles/variants/lib_g*.so); do
  v=$(basename $f .so)
  cp $f biscuit_b200/libbiscuit_b200.so
  timeout 200 python bench.py --tiles 4096 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> /tmp/err_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$v', 'tiles/s %.0f' % d['value'], ' '.join('%s=%.1f' % (n, k[n]['ms']) for n in ('sepconv_mid','maxpool_add','subsample','depthwise') if n in k))"
done
cp /tmp/lib_default.so biscuit_b200/libbiscuit_b200.so
