set -x
timeout 600 python -m pytest tests/test_loss_gpu.py -q -x -k "half_gram or anchor_shards or golden or same_operands or graphed" > gpurun_out/r03d_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r03d_loss.log
tail -n 4 gpurun_out/r03d_loss.log | cut -c1-300
python bench.py --workload c5_train --steps 5 > gpurun_out/r03d_c5.json 2> gpurun_out/r03d.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r03d_c1.json 2>> gpurun_out/r03d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"icl_fwd_sym" -s 2 -c 2 -o gpurun_out/r03d_fwdsym python scripts/profile_step.py c5_train > gpurun_out/r03d_ncu.log 2>&1
