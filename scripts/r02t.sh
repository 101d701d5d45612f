set -x
timeout 600 python -m pytest tests/test_loss_gpu.py -q -x > gpurun_out/r02t_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_loss.log
tail -n 3 gpurun_out/r02t_loss.log
python bench.py --workload c5_train --steps 5 > gpurun_out/r02t_c5.json 2> gpurun_out/r02t.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r02t_c1.json 2>> gpurun_out/r02t.err
tail -c 300 gpurun_out/r02t.err
