timeout 300 python scripts/gpu_gemm_epi.py > gpurun_out/gemm_epi2.log 2>&1; echo exit=$? >> gpurun_out/gemm_epi2.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu9.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu9.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.log 2>&1; echo exit=$? >> gpurun_out/bench11.log
cat gpurun_out/gemm_epi2.log; tail -3 gpurun_out/pytest_gpu9.log; tail -c 500 gpurun_out/bench11.log
