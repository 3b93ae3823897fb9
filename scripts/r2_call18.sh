set -x
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | tail -28
timeout 300 python scripts/gpu_proposals_time.py 64 2>&1 | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 1 -f -o gpurun_out/r2n_gemm_qkv_sam python scripts/profile_kernels.py gemm_qkv_sam 8 3 > gpurun_out/r2n_ncu_qkv.log 2>&1; echo exit=$?; tail -2 gpurun_out/r2n_ncu_qkv.log
