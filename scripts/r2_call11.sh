set -x
timeout 900 python -m pytest tests/test_gpu_proposals.py -q -m gpu -x -s 2>&1 | tail -30 | cut -c1-300
