# final round-2 evidence on one B200: tests, headline bench, T sweep, thresholding rows, launch list, ncu captures
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2_final.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'roof', d['roofline'], 'clocks', d['clocks'], 'cpu', d['cpu_baseline']); print(json.dumps(d['kernels']))"
tail -2 gpurun_out/bench_r2_final.err
for T in 10 100; do
  timeout 300 python bench.py --steps 3 --warmup 3 --T $T --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('T=$T tiles/s %.0f' % d['value'], 'head_fused ms %.2f head_gemm %.2f step ms %.1f' % (k['head_fused']['ms'], k['head_gemm']['ms'], d['ms_per_step']))"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; tail -c 700 gpurun_out/bench_r2_reference.json
timeout 600 python profiles/threshold_bench.py > gpurun_out/threshold_bench_r2.md 2> gpurun_out/threshold_bench_r2.err; cat gpurun_out/threshold_bench_r2.md | head -14
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2z.csv python profiles/run_predict.py 1024 512 > gpurun_out/launch_run.log 2>&1
tail -1 gpurun_out/launch_run.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv_mid -s 24 -c 3 -o gpurun_out/sepmid_r2z -f python profiles/run_predict.py 1024 512 > gpurun_out/sepmid_ncu.log 2>&1
tail -1 gpurun_out/sepmid_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv2d_fused -c 4 -o gpurun_out/sep2d_r2z -f python profiles/run_predict.py 512 512 > gpurun_out/sep2d_ncu.log 2>&1
tail -1 gpurun_out/sep2d_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 4 -c 1 -o gpurun_out/gemm_r2z -f python profiles/run_predict.py 512 512 > gpurun_out/gemm_ncu.log 2>&1
tail -1 gpurun_out/gemm_ncu.log
