set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm or norm_folded or statistics" 2>&1 | tail -8
timeout 600 python scripts/gpu_gemm_ab.py LLMSEG_GEMM_EPI 3 2>&1 | tail -14
