timeout 300 python scripts/gpu_hbm_check.py > gpurun_out/hbm3.log 2>&1; echo exit=$? >> gpurun_out/hbm3.log; tail -12 gpurun_out/hbm3.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu11.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu11.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.log 2>&1; echo exit=$? >> gpurun_out/bench13.log
tail -3 gpurun_out/pytest_gpu11.log; tail -c 600 gpurun_out/bench13.log
