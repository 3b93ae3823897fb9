set -x
for i in 1 2 3; do
  LLMSEG_B200_LIB=$PWD/llmseg_b200/libllmseg_b200_prev.so timeout 120 python scripts/gpu_qkv_time.py 2>&1 | tail -1
  timeout 120 python scripts/gpu_qkv_time.py 2>&1 | tail -1
done
