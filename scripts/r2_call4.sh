set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or relpos or window or select or small" 2>&1 | tail -15
timeout 300 python scripts/gpu_attn_time.py 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "index or reduced_depth_batched or sam_encoder" 2>&1 | grep -v "^$" | tail -60 | cut -c1-220
