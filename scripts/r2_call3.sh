set -x
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2c_pytest.log 2>&1; echo exit=$? >> gpurun_out/r2c_pytest.log
grep -n "max|d| per image\|index sweep\|50 calls\|sam encoder\|passed\|failed\|FAILED\|Error" gpurun_out/r2c_pytest.log | cut -c1-300 | head -80
sed -n '/== measured parity/,$p' gpurun_out/r2c_pytest.log | head -120
