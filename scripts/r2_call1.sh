set -x
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2a_pytest.log 2>&1; echo exit=$? >> gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_pytest.log
timeout 900 python tests/parity_bisect.py --seeds 4 --fold-norm 1 > gpurun_out/r2a_bisect_fold1.txt 2> gpurun_out/r2a_bisect_fold1.err; echo exit=$?
timeout 900 python tests/parity_bisect.py --seeds 4 --fold-norm 0 > gpurun_out/r2a_bisect_fold0.txt 2> gpurun_out/r2a_bisect_fold0.err; echo exit=$?
cat gpurun_out/r2a_bisect_fold1.txt; tail -5 gpurun_out/r2a_bisect_fold1.err
cat gpurun_out/r2a_bisect_fold0.txt; tail -5 gpurun_out/r2a_bisect_fold0.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'_kernel' -c 1800 --csv --log-file gpurun_out/r2a_launches_b1.csv python bench.py --batch 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_b1.log 2>&1; echo exit=$?; tail -2 gpurun_out/r2a_ncu_b1.log; wc -l gpurun_out/r2a_launches_b1.csv
