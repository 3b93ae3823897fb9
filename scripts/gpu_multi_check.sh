timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.log 2>&1; echo exit=$? >> gpurun_out/scale_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n2.log 2>&1; echo exit=$? >> gpurun_out/scale_n2.log
tail -c 600 gpurun_out/scale_n1.log; echo; tail -c 1500 gpurun_out/scale_n2.log
