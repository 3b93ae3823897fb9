# one GPU-box visit: validate the ping-pong attention kernel, then bench with it
LLMSEG_ATTN_V2=1 timeout 200 python scripts/gpu_attn_check.py > gpurun_out/attn5_v2.log 2>&1; rc=$?; echo exit=$rc >> gpurun_out/attn5_v2.log
if [ $rc -ne 0 ]; then tail -20 gpurun_out/attn5_v2.log; exit 1; fi
LLMSEG_ATTN_V2=1 timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu7_v2.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu7_v2.log
LLMSEG_ATTN_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7_v2.log 2>&1; echo exit=$? >> gpurun_out/bench7_v2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; echo exit=$? >> gpurun_out/bench7.log
grep -E "attention [0-9]|maxerr" gpurun_out/attn5_v2.log; tail -3 gpurun_out/pytest_gpu7_v2.log; tail -c 500 gpurun_out/bench7_v2.log; tail -c 300 gpurun_out/bench7.log
