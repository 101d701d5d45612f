# one-shot check of the templated prologue kernels: GPU suite without the oracle-bound / subprocess cases, then the loss slices
set -x
timeout 75 python -m pytest tests -m gpu -q -x --ignore=tests/test_eval_baseline_gpu.py --ignore=tests/test_reference_e2e_gpu.py > gpurun_out/r06c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r06c_pytest.log
tail -n 3 gpurun_out/r06c_pytest.log
timeout 25 python bench.py --workload c5_train --steps 10 > gpurun_out/r06c_c5.json 2> gpurun_out/r06c.err
timeout 20 python bench.py --workload c1_train --steps 10 > gpurun_out/r06c_c1.json 2>> gpurun_out/r06c.err
python - <<'PY'
import json
for f in ("gpurun_out/r06c_c5.json", "gpurun_out/r06c_c1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
