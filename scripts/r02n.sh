timeout 600 python scripts/onepass_probe.py 400000 1.5 > gpurun_out/r02n_probe.log 2>&1
timeout 900 python -m pytest tests/test_eval_gpu.py tests/test_seeds_gpu.py tests/test_mining_gpu.py -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -n 3 gpurun_out/r02n_pytest.log
timeout 900 python scripts/onepass_tune.py c4_1m 1.5,2.0 > gpurun_out/r02n_tune_1m.log 2>&1
