set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or relpos or window" 2>&1 | tail -5
timeout 300 python scripts/gpu_attn_time.py 2>&1 | head -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn4_kernel -s 1 -c 1 -f -o gpurun_out/r2g_attn4_global python scripts/profile_kernels.py attn_global 8 3 > gpurun_out/r2g_ncu_attn4.log 2>&1; echo exit=$?
