set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 > gpurun_out/r2t_bench_b8.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2t_bench_b8.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for g in d["roofline_all_gemms"]["groups"]: print(g)
print(d["roofline_attn"]["frac"], d["roofline_all_gemms"]["frac"])
PY
