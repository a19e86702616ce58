timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "injected or device_resident or micro_batch or repeatable or standardized" 2>&1 | tail -4
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f pageable %.0f' % (d['value'], d['e2e']['value'], d['e2e']['pageable_value']))"
