timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
python -c "
import json
d=json.load(open('gpurun_out/full_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']); print(json.dumps(d['kernels']))"
tail -3 gpurun_out/full_bench.err
