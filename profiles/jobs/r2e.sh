timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "stage_parity or fused_middle or features_and_uq or repeatable" 2>&1 | tail -6
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print(d['value'], d['ms_per_step']); print(json.dumps(d['kernels']))"
tail -3 gpurun_out/r2e_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv_mid -s 1 -c 3 -o gpurun_out/sepmid_r2e -f python profiles/run_predict.py 512 512 > gpurun_out/sepmid_ncu.log 2>&1
tail -2 gpurun_out/sepmid_ncu.log
