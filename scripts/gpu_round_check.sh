timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu17.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu17.log
tail -5 gpurun_out/pytest_gpu17.log
if ! grep -q "exit=0" gpurun_out/pytest_gpu17.log; then grep -n "Error\|assert\|FAILED" gpurun_out/pytest_gpu17.log | head -30; exit 1; fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench19.log 2>&1; echo exit=$? >> gpurun_out/bench19.log
timeout 600 python scripts/gpu_op_breakdown.py 8 > gpurun_out/opbreak_b8.log 2>&1; echo exit=$? >> gpurun_out/opbreak_b8.log
K='regex:add_rows_bcast_kernel|attn_kernel|attn_win_kernel|embed_splice_kernel|fill_kv_rows_kernel|gemm2_kernel|gemm_kernel|im2col3x3_kernel|maskpool_adjoint_kernel|maskpool_apply_kernel|maskpool_final_kernel|norm_kernel|patchify_kernel|select_kernel|small_attn_kernel|relpos_win_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1110 -c 555 --csv --log-file gpurun_out/launches_r01l.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo exit=$? >> gpurun_out/ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 1 -c 1 -f -o gpurun_out/prof_gemm_mlp1_r01l python scripts/profile_kernels.py gemm_mlp1 8 3 > gpurun_out/ncu_mlp1.log 2>&1; echo exit=$? >> gpurun_out/ncu_mlp1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 1 -c 1 -f -o gpurun_out/prof_attn_global_r01l python scripts/profile_kernels.py attn_global 8 3 > gpurun_out/ncu_attn.log 2>&1; echo exit=$? >> gpurun_out/ncu_attn.log
head -20 gpurun_out/opbreak_b8.log; tail -c 2500 gpurun_out/bench19.log; tail -2 gpurun_out/ncu_bench.log gpurun_out/ncu_mlp1.log gpurun_out/ncu_attn.log; wc -l gpurun_out/launches_r01l.csv
