timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r2d_tests.log; tail -8 gpurun_out/r2d_tests.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print(d['value'], d['ms_per_step']); print(json.dumps(d['kernels']))"
tail -3 gpurun_out/r2d_bench.err
