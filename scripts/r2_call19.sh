set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_proposals.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | tail -26
for kb in 32 16; do
  LLMSEG_GEMM_SK_MIN_KB=$kb timeout 300 python bench.py --batch 1 --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SK_MIN_KB=$kb batch1', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
LLMSEG_GEMM_SK_MIN_KB=32 timeout 300 python bench.py --batch 4 --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 | cut -c1-200
timeout 300 python bench.py --batch 4 --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | tail -1 | cut -c1-200
