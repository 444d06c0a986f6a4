timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2c_pytest.log
