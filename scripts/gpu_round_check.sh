# one GPU-box visit
timeout 300 python scripts/gpu_attn_check.py > gpurun_out/attn4.log 2>&1; echo exit=$? >> gpurun_out/attn4.log
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu6.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu6.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.log 2>&1; echo exit=$? >> gpurun_out/bench6.log
LLMSEG_GEMM_2CTA=1 timeout 150 python scripts/gpu_dev_check.py gemm time > gpurun_out/dev_gemm_2cta.log 2>&1; rc=$?; echo exit=$rc >> gpurun_out/dev_gemm_2cta.log
if [ $rc -eq 0 ]; then
  LLMSEG_GEMM_2CTA=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/pytest_gpu6_2cta.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu6_2cta.log
  LLMSEG_GEMM_2CTA=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_2cta.log 2>&1; echo exit=$? >> gpurun_out/bench6_2cta.log
fi
tail -3 gpurun_out/pytest_gpu6.log; tail -c 400 gpurun_out/bench6.log; tail -12 gpurun_out/dev_gemm_2cta.log
