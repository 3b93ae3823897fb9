set -x
timeout 600 python -m pytest tests/test_gpu_proposals.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | head -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel" -s 1 -c 1 -f -o gpurun_out/r2s_gemm_up1 python scripts/profile_kernels.py gemm_up1 8 3 > gpurun_out/r2s_ncu_gemm_up1.log 2>&1; echo exit=$?; tail -1 gpurun_out/r2s_ncu_gemm_up1.log
