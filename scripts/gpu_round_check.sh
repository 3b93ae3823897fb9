timeout 200 python scripts/gpu_attn_check.py win clip > gpurun_out/attn7_win.log 2>&1; rc=$?; echo exit=$rc >> gpurun_out/attn7_win.log
grep -E "attention [0-9]|maxerr|exit|Error" gpurun_out/attn7_win.log
if [ $rc -ne 0 ]; then tail -20 gpurun_out/attn7_win.log; exit 1; fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu8.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu8.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2>&1; echo exit=$? >> gpurun_out/bench9.log
tail -3 gpurun_out/pytest_gpu8.log; tail -c 600 gpurun_out/bench9.log
