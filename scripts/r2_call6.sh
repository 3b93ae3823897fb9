set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or relpos or window" 2>&1 | tail -5
timeout 300 python scripts/gpu_attn_time.py 2>&1 | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3_kernel -s 1 -c 1 -f -o gpurun_out/r2f_attn3_global python scripts/profile_kernels.py attn_global 8 3 > gpurun_out/r2f_ncu_attn3.log 2>&1; echo exit=$?
