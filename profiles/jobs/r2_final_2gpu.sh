timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_threshold_gpu.py -x -q -k "two_rank or sharded or native_comm" 2>&1 | tail -4
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_r2.json 2> gpurun_out/bench_2gpu_r2.err
echo "wall seconds for the N=2 bench (steps 20, warmup 3): $SECONDS"
python -c "
import json
d=json.loads(open('gpurun_out/bench_2gpu_r2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['cpu_baseline'], d['clocks'])"
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref_r2.json 2> gpurun_out/bench_2gpu_ref_r2.err
echo "wall seconds for the N=2 reference arm: $SECONDS"; tail -c 400 gpurun_out/bench_2gpu_ref_r2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload cohort --slides 64 --tiles 2000 --steps 2 --warmup 1 > gpurun_out/cohort_2gpu_r2.json 2> gpurun_out/cohort_2gpu_r2.err
python -c "
import json
d=json.loads(open('gpurun_out/cohort_2gpu_r2.json').read().strip().splitlines()[-1]); print('cohort', d['value'], d['slides_per_sec'], d['sharded_equals_single_process_apply'], d['ms_per_step'])"
