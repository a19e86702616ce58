timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "degenerate or stage_parity or injected or bench_micro or config1" 2>&1 | tail -15
