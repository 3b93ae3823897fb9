# one GPU-box visit: tests, bench, launch list
KREGEX='regex:gemm_kernel|attn_kernel|norm_kernel|patchify|im2col|embed_splice|add_rows|maskpool|small_attn|select_kernel'
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu5.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu5.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.log 2>&1; echo exit=$? >> gpurun_out/bench5.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 1396 -c 700 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo exit=$? >> gpurun_out/ncu_bench.log
tail -3 gpurun_out/pytest_gpu5.log; tail -c 700 gpurun_out/bench5.log; grep -c . gpurun_out/launches_b8.csv
