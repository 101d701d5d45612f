set -x
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --durations=8 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
python bench.py --workload c5_train --steps 5 > gpurun_out/r02f_c5.json 2> gpurun_out/r02f.err
python bench.py --workload c1_train --steps 10 > gpurun_out/r02f_c1.json 2>> gpurun_out/r02f.err
python bench.py --workload c1 --steps 10 --no-train --no-context > gpurun_out/r02f_c1_eval.json 2>> gpurun_out/r02f.err
python bench.py --workload c2 --steps 10 --no-train --no-context --no-audit > gpurun_out/r02f_c2_eval.json 2>> gpurun_out/r02f.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiIclFwd" -s 30 -c 2 -o gpurun_out/r02f_iclfwd python scripts/profile_step.py c5_train > gpurun_out/r02f_ncu_iclfwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"icl_bwd_fused" -s 1 -c 1 -o gpurun_out/r02f_fused python scripts/profile_step.py c5_train > gpurun_out/r02f_ncu_fused.log 2>&1
tail -8 gpurun_out/r02f_pytest.log
