for i in 1 2 3; do
  LLMSEG_B200_LIB=$PWD/llmseg_b200/libllmseg_b200_prev.so timeout 120 python scripts/gpu_attn_win_time.py 2>&1 | tail -1
  timeout 120 python scripts/gpu_attn_win_time.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "window or relpos" 2>&1 | tail -2
