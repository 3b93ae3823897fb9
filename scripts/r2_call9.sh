set -x
timeout 300 python scripts/gpu_attn_time.py 2>&1 | head -12
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo exit=$? >> gpurun_out/r2h_pytest.log; tail -5 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo exit=$?; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','batch1','gpu_launches_per_step')}); print(d['roofline_attn']['achieved'], d['roofline_attn']['frac'], d['roofline_all_gemms'])"
