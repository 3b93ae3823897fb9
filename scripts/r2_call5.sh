set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3_kernel -s 1 -c 1 -f -o gpurun_out/r2e_attn3_global python scripts/profile_kernels.py attn_global 8 3 > gpurun_out/r2e_ncu_attn3.log 2>&1; echo exit=$?; tail -3 gpurun_out/r2e_ncu_attn3.log
ls -la gpurun_out/*.ncu-rep | tail -3
