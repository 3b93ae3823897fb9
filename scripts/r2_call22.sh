set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "v5_lagged or relpos_attention" 2>&1 | grep -v "^$" | tail -12
timeout 300 python scripts/gpu_attn_time.py global 2>&1 | tail -10
timeout 600 python -m pytest tests/test_gpu_proposals.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | grep "wall\|mask_stats"
