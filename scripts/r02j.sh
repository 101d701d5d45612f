set -x
timeout 600 python scripts/onepass_probe.py 400000 1.5 > gpurun_out/r02j_probe.log 2>&1
timeout 900 python scripts/onepass_tune.py c4_1m 1.5,2.0,2.5 > gpurun_out/r02j_tune_1m.log 2>&1
timeout 900 python -m pytest tests/test_eval_baseline_gpu.py -q -x -s -k "two_sweep_size or one_pass" > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_mining_gpu.py tests/test_seeds_gpu.py -q -x > gpurun_out/r02j_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest2.log
tail -n 5 gpurun_out/r02j_pytest.log gpurun_out/r02j_pytest2.log
