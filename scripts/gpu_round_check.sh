timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu15.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu15.log
tail -5 gpurun_out/pytest_gpu15.log
if ! grep -q "exit=0" gpurun_out/pytest_gpu15.log; then exit 1; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.log 2>&1; echo exit=$? >> gpurun_out/bench17.log
timeout 600 python scripts/gpu_op_breakdown.py 8 > gpurun_out/opbreak_b8.log 2>&1; echo exit=$? >> gpurun_out/opbreak_b8.log
timeout 600 python scripts/gpu_gemm_epi.py > gpurun_out/gemm_epi3.log 2>&1; echo exit=$? >> gpurun_out/gemm_epi3.log
head -26 gpurun_out/opbreak_b8.log; cat gpurun_out/gemm_epi3.log; tail -c 700 gpurun_out/bench17.log
