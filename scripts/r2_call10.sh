set -x
timeout 900 python -m pytest tests/test_gpu_proposals.py -q -m gpu -x -s 2>&1 | tail -40 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -k "full_depth_batch8 or constructor" 2>&1 | tail -5
