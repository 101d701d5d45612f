set -x
timeout 900 python -m pytest tests/test_eval_gpu.py tests/test_mining_gpu.py tests/test_seeds_gpu.py -q -x > gpurun_out/r03i_eval.log 2>&1; echo "rc=$?" >> gpurun_out/r03i_eval.log
tail -n 3 gpurun_out/r03i_eval.log | cut -c1-200
timeout 600 python -m pytest tests/test_eval_baseline_gpu.py -q -x -k "baseline_config" > gpurun_out/r03i_base.log 2>&1; echo "rc=$?" >> gpurun_out/r03i_base.log
tail -n 3 gpurun_out/r03i_base.log | cut -c1-200
python scripts/profile_eval.py > gpurun_out/r03i_eval_profile_1m.txt 2>/dev/null
head -n 12 gpurun_out/r03i_eval_profile_1m.txt | cut -c1-150
