timeout 300 python scripts/gpu_attn_time.py > gpurun_out/attn_time.log 2>&1
LLMSEG_ATTN_V2=1 timeout 300 python scripts/gpu_attn_time.py >> gpurun_out/attn_time.log 2>&1
LLMSEG_ATTN_V2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" > gpurun_out/pytest_v2.log 2>&1; echo exit=$? >> gpurun_out/pytest_v2.log
timeout 600 python scripts/gpu_gemm_epi.py > gpurun_out/gemm_epi5.log 2>&1; echo exit=$? >> gpurun_out/gemm_epi5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_win_kernel -s 1 -c 1 -f -o gpurun_out/prof_attn_win_r01k python scripts/profile_kernels.py attn_window 8 3 > gpurun_out/ncu_win.log 2>&1; echo exit=$? >> gpurun_out/ncu_win.log
cat gpurun_out/attn_time.log; tail -3 gpurun_out/pytest_v2.log; tail -8 gpurun_out/gemm_epi5.log; tail -2 gpurun_out/ncu_win.log
