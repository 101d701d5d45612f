# ncu launch lists (one gpu__time_duration pass, no clock control) of the loss-layer slice of the final build
set -x
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r06a_launches_c5_train.csv python bench.py --workload c5_train --steps 1 --warmup 3 --no-graph > gpurun_out/r06a_ncu_bench_c5.log 2>&1
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r06a_launches_c1_train.csv python bench.py --workload c1_train --steps 1 --warmup 3 --no-graph > gpurun_out/r06a_ncu_bench_c1.log 2>&1
ls -la gpurun_out | tail -n 6
