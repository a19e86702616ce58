# BASELINE config 4: MC-dropout sample sweep T = 10 / 30 / 100 on 100 k tiles (one backbone pass, T head passes)
for T in 10 30 100; do
  timeout 400 python bench.py --tiles 100000 --steps 1 --warmup 3 --T $T --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
tot=sum(v['ms'] for v in k.values())
print('T=$T  tiles/s %.0f  step %.1f ms  head_fused %.1f ms + head_gemm %.1f ms = %.2f %% of the kernel time' % (d['value'], d['ms_per_step'], k['head_fused']['ms'], k['head_gemm']['ms'], 100*(k['head_fused']['ms']+k['head_gemm']['ms'])/tot))"
done
