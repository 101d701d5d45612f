for v in opx5 opx1 opx2 opx4; do
  echo "== $v" >> gpurun_out/r02o_probe.log
  SNAG_B200_LIB=$PWD/snag_b200/_variants/lib_$v.so timeout 300 python scripts/onepass_probe.py 400000 1.5 2>&1 | grep eval_onepass | tail -n 1 >> gpurun_out/r02o_probe.log
done
echo "== base" >> gpurun_out/r02o_probe.log
timeout 300 python scripts/onepass_probe.py 400000 1.5 2>&1 | grep eval_onepass | tail -n 1 >> gpurun_out/r02o_probe.log
