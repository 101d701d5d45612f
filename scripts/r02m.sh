timeout 600 python scripts/onepass_probe.py 400000 1.5 > gpurun_out/r02m_probe.log 2>&1
timeout 600 python -m pytest tests/test_eval_gpu.py -q -x -k "two_sweep or oracle" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -n 3 gpurun_out/r02m_pytest.log
