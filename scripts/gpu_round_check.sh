LLMSEG_ATTN_V2=1 timeout 200 python scripts/gpu_attn_check.py > gpurun_out/attn6_v2.log 2>&1; rc=$?; echo exit=$rc >> gpurun_out/attn6_v2.log
grep -E "attention [0-9]|maxerr|exit" gpurun_out/attn6_v2.log
if [ $rc -ne 0 ]; then tail -20 gpurun_out/attn6_v2.log; exit 1; fi
LLMSEG_ATTN_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_v2.log 2>&1; echo exit=$? >> gpurun_out/bench8_v2.log
tail -c 600 gpurun_out/bench8_v2.log
