timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "dropout_site or injected or philox or fused_head or sample_sweep" 2>&1 | tail -8
