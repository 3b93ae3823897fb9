set -x
timeout 900 python -m pytest tests/test_gpu_proposals.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | head -6
