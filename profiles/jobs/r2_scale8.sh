SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_r2.json 2> gpurun_out/bench_8gpu_r2.err
echo "wall seconds: $SECONDS"
python -c "
import json
d=json.loads(open('gpurun_out/bench_8gpu_r2.json').read().strip().splitlines()[-1]); print('N=8', d['value'], d['e2e'], d['ms_per_step'], d['clocks'])"
tail -2 gpurun_out/bench_8gpu_r2.err
