timeout 300 python scripts/gpu_hbm_check.py > gpurun_out/hbm2.log 2>&1; echo exit=$? >> gpurun_out/hbm2.log; cat gpurun_out/hbm2.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu10.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu10.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.log 2>&1; echo exit=$? >> gpurun_out/bench12.log
tail -3 gpurun_out/pytest_gpu10.log; tail -c 500 gpurun_out/bench12.log
