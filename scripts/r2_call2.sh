set -x
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/r2b_pytest.log 2>&1; echo exit=$? >> gpurun_out/r2b_pytest.log; tail -80 gpurun_out/r2b_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python tests/parity_bisect.py --seeds 4 --fold-norm image > gpurun_out/r2b_bisect_image.txt 2> gpurun_out/r2b_bisect_image.err; echo exit=$?
cat gpurun_out/r2b_bisect_image.txt; tail -5 gpurun_out/r2b_bisect_image.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo exit=$?; cat gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
