set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^attn_kernel" -s 1 -c 1 -f -o gpurun_out/r2w_attn_global python scripts/profile_kernels.py attn_global 8 3 > gpurun_out/r2w_ncu_ag.log 2>&1; echo exit=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_win_kernel" -s 1 -c 1 -f -o gpurun_out/r2w_attn_window python scripts/profile_kernels.py attn_window 8 3 > gpurun_out/r2w_ncu_aw.log 2>&1; echo exit=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 1 -f -o gpurun_out/r2w_gemm_qkv_sam python scripts/profile_kernels.py gemm_qkv_sam 8 3 > gpurun_out/r2w_ncu_qkv.log 2>&1; echo exit=$?
ls -la gpurun_out/r2w_*.ncu-rep
