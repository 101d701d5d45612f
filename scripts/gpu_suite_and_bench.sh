set -x
timeout 2400 python -m pytest tests -m gpu -q --maxfail=20 --durations=8 > gpurun_out/r05c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r05c_pytest.log
tail -n 6 gpurun_out/r05c_pytest.log
timeout 1200 python bench.py --steps 3 > gpurun_out/r05c_bench.json 2> gpurun_out/r05c_bench.err
tail -c 300 gpurun_out/r05c_bench.err
