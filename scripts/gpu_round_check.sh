timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu16.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu16.log
tail -5 gpurun_out/pytest_gpu16.log
if ! grep -q "exit=0" gpurun_out/pytest_gpu16.log; then grep -n "Error\|assert\|FAILED" gpurun_out/pytest_gpu16.log | head -30; exit 1; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench18.log 2>&1; echo exit=$? >> gpurun_out/bench18.log
timeout 600 python scripts/gpu_op_breakdown.py 8 > gpurun_out/opbreak_b8.log 2>&1; echo exit=$? >> gpurun_out/opbreak_b8.log
timeout 600 python scripts/gpu_gemm_epi.py > gpurun_out/gemm_epi4.log 2>&1; echo exit=$? >> gpurun_out/gemm_epi4.log
LLMSEG_GEMM_TMA_STORE=0 timeout 600 python scripts/gpu_gemm_epi.py > gpurun_out/gemm_epi4_direct.log 2>&1; echo exit=$? >> gpurun_out/gemm_epi4_direct.log
head -24 gpurun_out/opbreak_b8.log; echo TMA; head -12 gpurun_out/gemm_epi4.log; echo DIRECT; head -12 gpurun_out/gemm_epi4_direct.log; tail -c 700 gpurun_out/bench18.log
