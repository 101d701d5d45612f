timeout 600 python scripts/onepass_power.py 400000 12 > gpurun_out/r02p_power.log 2>&1
