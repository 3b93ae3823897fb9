set -x
timeout 900 python -m pytest tests/test_gpu_proposals.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | tail -26
LLMSEG_AMG_FUSED_UPSCALE=0 timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | grep "wall"
