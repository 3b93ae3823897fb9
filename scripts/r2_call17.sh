set -x
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2m_pytest.log 2>&1; echo exit=$? >> gpurun_out/r2m_pytest.log; tail -4 gpurun_out/r2m_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
LLMSEG_FOLD_NORM=0 timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "full_depth_dinov2" 2>&1 | grep "dinov2 pred" 
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo exit=$?; tail -3 gpurun_out/r2m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json'))
for k in ('value','ms_per_step','e2e','batch1','configs3','configs4','proposals','eager_gpu_baseline','cpu_baseline','roofline','clocks'): print(k, d.get(k))
print('gemms', {k: v for k, v in d['roofline_all_gemms'].items() if k != 'groups'})
for g in d['roofline_all_gemms']['groups']: print('  ', g)
print('attn', d['roofline_attn'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_reference.json 2>/dev/null; cat gpurun_out/r2m_reference.json | cut -c1-700
