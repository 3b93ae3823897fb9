set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_scale_n2.json 2> gpurun_out/r2j_scale_n2.err; echo exit=$?
tail -3 gpurun_out/r2j_scale_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2j_scale_n2.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','e2e','n_gpus','configs3','configs4','gather_check','per_rank')})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-extra-configs --sync-gather > gpurun_out/r2j_scale_n2_sync.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2j_scale_n2_sync.json').read().strip().splitlines()[-1]); print('sync gather', {k: d[k] for k in ('value','ms_per_step','per_rank')})"
