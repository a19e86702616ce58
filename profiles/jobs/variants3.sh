cp biscuit_b200/libbiscuit_b200.so /tmp/lib_default.so
run() {
  v=$(basename $1 .so); cp $1 biscuit_b200/libbiscuit_b200.so
  timeout 200 python bench.py --tiles 4024 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> /tmp/err_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$v', 'tiles/s %.0f' % d['value'], ' '.join('%s=%.1f' % (n, k[n]['ms']) for n in ('sepconv_mid','gemm_pointwise','sepconv_fused') if n in k))"
  tail -2 /tmp/err_$v.log | cut -c1-300
}
for f in profiles/variants/lib_*.so; do run $f; done
for f in $(ls -r profiles/variants/lib_*.so); do run $f; done
if [ -n "$CHECK" ]; then
  cp profiles/variants/lib_$CHECK.so biscuit_b200/libbiscuit_b200.so
  timeout 400 python -m pytest tests/test_model_gpu.py -x -q -k "$CHECKK" 2>&1 | tail -3
fi
cp /tmp/lib_default.so biscuit_b200/libbiscuit_b200.so
