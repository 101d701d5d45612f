set -x
timeout 900 python -m pytest tests/test_eval_gpu.py -q -x -k "large_k or dropin or materialised" > gpurun_out/r02w_eval.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_eval.log
tail -n 5 gpurun_out/r02w_eval.log
timeout 600 python -m pytest tests/test_eval_baseline_gpu.py -q -x -k "streamed" > gpurun_out/r02w_streamed.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_streamed.log
tail -n 3 gpurun_out/r02w_streamed.log
timeout 600 python -m pytest tests/test_loss_gpu.py -q -x > gpurun_out/r02w_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_loss.log
tail -n 3 gpurun_out/r02w_loss.log
timeout 600 python scripts/e2e_probe.py > gpurun_out/r02w_e2e_probe.log 2>&1
grep -v Warn gpurun_out/r02w_e2e_probe.log | tail -n 12
