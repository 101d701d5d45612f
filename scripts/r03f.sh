set -x
timeout 900 python -m pytest tests/test_loss_gpu.py -q -x > gpurun_out/r03f_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r03f_loss.log
tail -n 12 gpurun_out/r03f_loss.log | cut -c1-300
python bench.py --workload c5_train --steps 5 > gpurun_out/r03f_c5.json 2> gpurun_out/r03f.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r03f_c1.json 2>> gpurun_out/r03f.err
