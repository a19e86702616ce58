timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "stage_parity or fused_middle or features_and_uq" 2>&1 | tail -25 > gpurun_out/r2b_tests.log; tail -25 gpurun_out/r2b_tests.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print(d['value'], d['ms_per_step']); print(json.dumps(d['kernels']))"
tail -3 gpurun_out/r2b_bench.err
