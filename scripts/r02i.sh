set -x
timeout 600 python scripts/onepass_probe.py 400000 1.5 > gpurun_out/r02i_probe.log 2>&1
tail -n 20 gpurun_out/r02i_probe.log
