timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'roof', d['roofline'], 'clocks', d['clocks'], 'cpu', d['cpu_baseline']); print(json.dumps(d['kernels']))"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2z.csv python profiles/run_predict.py 1024 512 > gpurun_out/launch_run.log 2>&1
tail -1 gpurun_out/launch_run.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv_mid -s 24 -c 3 -o gpurun_out/sepmid_r2z -f python profiles/run_predict.py 1024 512 > gpurun_out/sepmid_ncu.log 2>&1
tail -1 gpurun_out/sepmid_ncu.log
