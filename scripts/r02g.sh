set -x
timeout 900 python -m pytest tests/test_eval_baseline_gpu.py -q -x -s -k "two_sweep_size or one_pass" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
timeout 300 python -m pytest tests/test_eval_gpu.py -q -x > gpurun_out/r02g_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest2.log
timeout 900 python bench.py --workload c4_1m --steps 2 --no-train --no-context > gpurun_out/r02g_1m.json 2> gpurun_out/r02g_1m.err
timeout 300 python bench.py --workload c4_100k --steps 3 --no-train --no-context --no-audit > gpurun_out/r02g_100k.json 2> gpurun_out/r02g_100k.err
tail -5 gpurun_out/r02g_pytest.log gpurun_out/r02g_pytest2.log
