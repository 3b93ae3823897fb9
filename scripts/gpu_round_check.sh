# investigation: where does the SAM global attention tile time go?
./tests/probe/mufu_bench > gpurun_out/mufu.log 2>&1
for dbg in 0 1 2 3 4; do
  echo "=== LLMSEG_ATTN_DBG=$dbg" >> gpurun_out/attn_knobs.log
  LLMSEG_ATTN_V2=1 LLMSEG_ATTN_DBG=$dbg timeout 120 python scripts/gpu_attn_check.py glob 2>&1 | grep -E "attention [0-9]" >> gpurun_out/attn_knobs.log
done
for dbg in 0 1; do
  echo "=== v1 LLMSEG_ATTN_DBG=$dbg" >> gpurun_out/attn_knobs.log
  LLMSEG_ATTN_DBG=$dbg timeout 120 python scripts/gpu_attn_check.py glob 2>&1 | grep -E "attention [0-9]" >> gpurun_out/attn_knobs.log
done
cat gpurun_out/mufu.log gpurun_out/attn_knobs.log
