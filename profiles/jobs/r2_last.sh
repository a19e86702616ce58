timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['pageable_value'], 'frac', d['roofline']['frac'], d['model_tensor_frac_of_sustained_peak'], d['gpu_launches'], d['config']['max_batch'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
