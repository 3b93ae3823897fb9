set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_proposals.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | tail -26
LLMSEG_T2I_V1=1 timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | grep "tok2img\|wall"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 > gpurun_out/r2o_bench_b8.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2o_bench_b8.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for g in d["roofline_all_gemms"]["groups"]: print(g)
print(d["roofline_attn"]["frac"], d["roofline_all_gemms"]["frac"])
PY
