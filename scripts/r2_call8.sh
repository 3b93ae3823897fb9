set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or relpos or window" 2>&1 | tail -5
timeout 300 python scripts/gpu_attn_time.py 2>&1 | head -14
