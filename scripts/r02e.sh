set -x
timeout 600 python -m pytest tests/test_reference_e2e_gpu.py -q > gpurun_out/r02e_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_e2e.log
python scripts/profile_step.py c5_train > gpurun_out/r02e_step_profile_c5.txt 2> gpurun_out/r02e_prof.err
python scripts/profile_step.py c1_train > gpurun_out/r02e_step_profile_c1.txt 2>> gpurun_out/r02e_prof.err
# launch list of one eager c5 step sequence (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file gpurun_out/r02e_launches_c5_train.csv python bench.py --workload c5_train --steps 1 --warmup 3 --no-graph > gpurun_out/r02e_ncu_bench.log 2>&1
# full captures: fused backward, ICL forward, the two wide-table backward kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"icl_bwd_fused" -s 1 -c 1 -o gpurun_out/r02e_fused python scripts/profile_step.py c5_train > gpurun_out/r02e_ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"EpiIclFwd" -s 30 -c 2 -o gpurun_out/r02e_iclfwd python scripts/profile_step.py c5_train > gpurun_out/r02e_ncu_iclfwd.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"noise_mask|col_stats|joint_fuse_fwd|prep_bf16|gauss_fill" -c 12 -o gpurun_out/r02e_bw python scripts/bench_rows.py --shape c1 > gpurun_out/r02e_rows_c1.jsonl 2> gpurun_out/r02e_rows.err
ls -la gpurun_out | tail -20
