set -x
timeout 600 python -m pytest tests/test_loss_gpu.py -q -k "fused" --maxfail=20 > gpurun_out/r02c_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/r02c_fused.log
if grep -q "fused rc=0" gpurun_out/r02c_fused.log; then
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --durations=8 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
  python bench.py --workload c5_train --steps 5 > gpurun_out/r02c_c5.json 2> gpurun_out/r02c_c5.err
  python bench.py --workload c1_train --steps 10 > gpurun_out/r02c_c1.json 2>> gpurun_out/r02c_c5.err
  SNAG_FUSED_BACKWARD=0 python bench.py --workload c5_train --steps 5 > gpurun_out/r02c_c5_unfused.json 2>> gpurun_out/r02c_c5.err
else
  SNAG_FUSED_BACKWARD=0 timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --durations=8 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest(unfused) rc=$?" >> gpurun_out/r02c_pytest.log
fi
tail -15 gpurun_out/r02c_fused.log; tail -12 gpurun_out/r02c_pytest.log
