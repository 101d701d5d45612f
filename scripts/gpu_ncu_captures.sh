set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03j_launches_c4_1m.csv python bench.py --steps 1 --warmup 3 --no-train --no-context --no-audit > gpurun_out/r03j_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03j_launches_c5_train.csv python bench.py --workload c5_train --steps 1 --warmup 3 --no-graph > gpurun_out/r03j_ncu_bench_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03j_launches_c1_train.csv python bench.py --workload c1_train --steps 1 --warmup 3 --no-graph > gpurun_out/r03j_ncu_bench_c1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"icl_bwd_fused" -s 1 -c 1 -o gpurun_out/r03j_fused python scripts/profile_step.py c5_train > gpurun_out/r03j_ncu_fused.log 2>&1
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section WarpStateStats --section Occupancy --clock-control none --kernel-name-base demangled -k regex:"EpiRowColTopKT" -s 1 -c 1 -o gpurun_out/r03j_onepass python scripts/profile_eval.py > gpurun_out/r03j_ncu_onepass.log 2>&1
ls -la gpurun_out | tail -n 12
