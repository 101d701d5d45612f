# final-build check: default bench line (stdout must be the JSON line alone) + the GPU suite without its slowest, oracle-bound case
set -x
timeout 170 python bench.py --steps 3 > gpurun_out/r06b_bench.json 2> gpurun_out/r06b_bench.err; echo "bench rc=$?"
wc -l gpurun_out/r06b_bench.json
timeout 120 python -m pytest tests -m gpu -q --maxfail=20 --durations=5 -k "not test_two_sweep_size_against_oracle" > gpurun_out/r06b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r06b_pytest.log
tail -n 4 gpurun_out/r06b_pytest.log
