timeout 900 python scripts/onepass_tune.py c4_1m 1.0,1.25,1.5,2.0 > gpurun_out/r02q_tune_1m.log 2>&1
timeout 900 python -m pytest tests/test_eval_baseline_gpu.py -q -x -s -k "two_sweep_size or one_pass" > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log
tail -n 4 gpurun_out/r02q_pytest.log
