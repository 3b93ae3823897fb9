KREGEX='regex:gemm_kernel|gemm2_kernel|attn_kernel|attn_win_kernel|norm_kernel|patchify|im2col|embed_splice|add_rows|maskpool|small_attn|select_kernel'
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.log 2>&1; echo exit=$? >> gpurun_out/bench10.log
LLMSEG_ATTN_WIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench10_nowin.log 2>&1; echo exit=$? >> gpurun_out/bench10_nowin.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 1396 -c 700 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo exit=$? >> gpurun_out/ncu_bench.log
tail -c 400 gpurun_out/bench10.log; tail -c 400 gpurun_out/bench10_nowin.log; grep -c . gpurun_out/launches_b8.csv
