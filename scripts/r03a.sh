set -x
timeout 600 python -m pytest tests/test_loss_gpu.py -q -x > gpurun_out/r03a_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r03a_loss.log
tail -n 25 gpurun_out/r03a_loss.log | cut -c1-400
python bench.py --workload c5_train --steps 5 > gpurun_out/r03a_c5.json 2> gpurun_out/r03a.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r03a_c1.json 2>> gpurun_out/r03a.err
SNAG_SYM_FORWARD=0 python bench.py --workload c1_train --steps 10 > gpurun_out/r03a_c1_nosym.json 2>> gpurun_out/r03a.err
grep -v Warn gpurun_out/r03a.err | tail -n 5
