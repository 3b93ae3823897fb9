set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | head -9
