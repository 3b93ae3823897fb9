timeout 300 python scripts/gpu_gemm_epi.py > gpurun_out/gemm_epi.log 2>&1; echo exit=$? >> gpurun_out/gemm_epi.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 1 -o gpurun_out/prof_gemm2_mlp1 python scripts/profile_kernels.py gemm_mlp1 8 2 > gpurun_out/ncu_gemm2_mlp1.log 2>&1
cat gpurun_out/gemm_epi.log
