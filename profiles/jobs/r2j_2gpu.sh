timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_threshold_gpu.py -x -q -k "two_rank or sharded or native_comm" 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload cohort --slides 64 --tiles 2000 --steps 2 --warmup 1 > gpurun_out/cohort_2gpu_r2.json 2> gpurun_out/cohort_2gpu_r2.err
tail -c 1500 gpurun_out/cohort_2gpu_r2.json; tail -3 gpurun_out/cohort_2gpu_r2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_r2.json 2> gpurun_out/bench_2gpu_r2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu_r2.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline'])"
