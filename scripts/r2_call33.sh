set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2y_pytest_gpu.log; cat gpurun_out/r2y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r2y_smoke.log; cat gpurun_out/r2y_smoke.log
SECONDS=0
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2y_bench_b8.json; echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2y_bench_b8.json"))
for k in ("value", "ms_per_step", "e2e", "clocks", "gpu_launches", "batch1", "proposals"):
    print(k, json.dumps(d.get(k))[:260])
print(d["roofline"]["frac"], d["roofline"].get("frac_of_burst"), d["roofline_attn"]["frac"], d["roofline_attn"].get("tensor_pipe_pct"), d["roofline_all_gemms"]["frac"])
PY
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2y_reference_arm.json; echo "reference arm wall ${SECONDS}s"; cut -c1-200 gpurun_out/r2y_reference_arm.json
