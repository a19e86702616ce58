timeout 600 python -m pytest tests/test_model_gpu.py -x -q -k "stage_parity or fused_middle or features_and_uq or repeatable" 2>&1 | tail -6
bash profiles/jobs/variants.sh
