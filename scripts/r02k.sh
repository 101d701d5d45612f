timeout 600 python scripts/onepass_probe.py 400000 1.5 > gpurun_out/r02k_probe.log 2>&1
