# quick safety check of the fused middle-flow kernel, then the model suite and a bench
timeout 240 python -m pytest tests/test_model_gpu.py -x -q -k "fused_middle_flow or stage_parity" 2>&1 | tail -6 || exit 1
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python -c "
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else None); print(' '.join('%s=%.1f' % (k, v['ms']) for k, v in d['kernels'].items()))"
tail -3 gpurun_out/quick_bench.err
