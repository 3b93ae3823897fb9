# one GPU-box visit: correctness of the new kernels, tests, bench, launch list
for cl in 1 2 4; do echo "=== LLMSEG_GEMM_CLUSTER=$cl" >> gpurun_out/dev_gemm.log; LLMSEG_GEMM_CLUSTER=$cl timeout 300 python scripts/gpu_dev_check.py gemm time >> gpurun_out/dev_gemm.log 2>&1; done
timeout 300 python scripts/gpu_attn_check.py > gpurun_out/attn2.log 2>&1; echo exit=$? >> gpurun_out/attn2.log
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu3.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu3.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.log 2>&1; echo exit=$? >> gpurun_out/bench3.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench3_b1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:llmseg -s 1396 -c 700 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo exit=$? >> gpurun_out/ncu_bench.log
tail -3 gpurun_out/pytest_gpu3.log; tail -c 700 gpurun_out/bench3.log; grep -c . gpurun_out/launches_b8.csv
