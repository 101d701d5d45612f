set -x
timeout 900 python -m pytest tests/test_loss_gpu.py -q -x > gpurun_out/r02s_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_loss.log
tail -n 4 gpurun_out/r02s_loss.log
timeout 600 python -m pytest tests/test_eval_baseline_gpu.py -q -x -k "streamed" > gpurun_out/r02s_streamed.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_streamed.log
tail -n 4 gpurun_out/r02s_streamed.log
timeout 300 python scripts/ial_diag.py > gpurun_out/r02s_ial_diag.log 2>&1
cat gpurun_out/r02s_ial_diag.log | tail -n 6
python bench.py --workload c5_train --steps 5 > gpurun_out/r02s_c5.json 2> gpurun_out/r02s.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r02s_c1.json 2>> gpurun_out/r02s.err
timeout 900 python bench.py --steps 3 --no-train --no-context --no-audit > gpurun_out/r02s_1m.json 2>> gpurun_out/r02s.err
tail -c 400 gpurun_out/r02s.err
