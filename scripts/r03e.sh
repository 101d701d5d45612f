set -x
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_reference_e2e_gpu.py -q -x > gpurun_out/r03e_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r03e_loss.log
tail -n 12 gpurun_out/r03e_loss.log | cut -c1-300
python bench.py --workload c5_train --steps 5 > gpurun_out/r03e_c5.json 2> gpurun_out/r03e.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r03e_c1.json 2>> gpurun_out/r03e.err
python scripts/profile_step.py c1_train > gpurun_out/r03e_step_profile_c1.txt 2>/dev/null
python scripts/profile_step.py c5_train > gpurun_out/r03e_step_profile_c5.txt 2>/dev/null
