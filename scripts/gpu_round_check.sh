# one GPU-box visit: correctness of the new kernels, tests, bench, launch list, ncu of the top kernel
KREGEX='regex:gemm_kernel|attn_kernel|norm_kernel|patchify|im2col|embed_splice|add_rows|maskpool|small_attn|select_kernel'
timeout 300 python scripts/gpu_attn_check.py > gpurun_out/attn3.log 2>&1; echo exit=$? >> gpurun_out/attn3.log
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu4.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu4.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4.log 2>&1; echo exit=$? >> gpurun_out/bench4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 1396 -c 700 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo exit=$? >> gpurun_out/ncu_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 1 -c 1 -o gpurun_out/prof_attn_global_b python scripts/profile_kernels.py attn_global 8 2 > gpurun_out/ncu_attn_global.log 2>&1
tail -3 gpurun_out/pytest_gpu4.log; tail -c 700 gpurun_out/bench4.log; grep -c . gpurun_out/launches_b8.csv
