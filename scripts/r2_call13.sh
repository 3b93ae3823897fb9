set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm or qkv or norm_folded or statistics or window_partition" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo exit=$?; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','batch1','gpu_launches_per_step','clocks')}); print(d['roofline']); print(d['roofline_attn']['achieved'], d['roofline_attn']['frac'], d['roofline_all_gemms'])"
LLMSEG_GEMM_EPI=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2i_bench_epi0.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_epi0.json')); print('EPI=0', {k: d[k] for k in ('value','ms_per_step')}, d['roofline_all_gemms'])"
