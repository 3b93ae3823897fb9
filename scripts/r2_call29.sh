set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2u_pytest_gpu.log; cat gpurun_out/r2u_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2u_bench_b8.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2u_bench_b8.json"))
for k in ("value", "ms_per_step", "e2e", "clocks", "gpu_launches", "batch1", "configs3", "configs4", "proposals"):
    print(k, json.dumps(d.get(k))[:300])
print(d["roofline"]["frac"], d["roofline"].get("frac_of_burst"), d["roofline_attn"]["frac"], d["roofline_all_gemms"]["frac"])
PY
