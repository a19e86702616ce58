# BASELINE config 3 at full cardinality: 1000 slides x 2000 tiles sharded over the ranks by whole slides
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --workload cohort --slides 1000 --tiles 2000 --steps 1 --warmup 1 > gpurun_out/cohort_full_2gpu_r2.json 2> gpurun_out/cohort_full_2gpu_r2.err
echo "wall seconds: $SECONDS"
python -c "
import json
d=json.loads(open('gpurun_out/cohort_full_2gpu_r2.json').read().strip().splitlines()[-1]); print('cohort', d['value'], d['slides_per_sec'], d['sharded_equals_single_process_apply'], d['ms_per_step'], d['config']['workload'][:120], d['apply_results'])"
tail -3 gpurun_out/cohort_full_2gpu_r2.err
