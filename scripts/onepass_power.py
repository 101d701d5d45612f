"""Dev probe: NVML clock / power / throttle reasons while the evaluation runs back to back (sustained), per mode."""
import json
import sys
import threading

import pynvml
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from snag_b200 import evaluate, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
d, k, sigma = 1200, 10, 6.5
dev = torch.device("cuda", 0)
emb, left, right = bench.synth_tables(n, d, sigma, dev)
X, xn = ops.prep_bf16(emb, left, True)
Y, yn = ops.prep_bf16(emb, right, True)
del emb
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
evaluate.ONE_PASS_GAMMA = 1.5
for mode in ("two_sweep", "one_pass", "two_sweep", "one_pass"):
    one = mode == "one_pass"
    evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=one)
    torch.cuda.synchronize()
    clk, pw, rs, tmp, stop = [], [], [], [], threading.Event()

    def sample():
        while not stop.is_set():
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000)
            rs.append(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
            tmp.append(pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU))
            stop.wait(0.05)
    th = threading.Thread(target=sample)
    th.start()
    ev = []
    ops.SWEEP_EVENT_SINK = ev
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=one)
    e1.record()
    torch.cuda.synchronize()
    ops.SWEEP_EVENT_SINK = None
    stop.set()
    th.join()
    per = {}
    for nm, a, b, *_ in ev:
        per.setdefault(nm, []).append(a.elapsed_time(b))
    half = len(clk) // 3
    reasons = {}
    for r in rs[half:]:
        reasons[hex(r)] = reasons.get(hex(r), 0) + 1
    print(json.dumps({"mode": mode, "ms_per_eval": round(e0.elapsed_time(e1) / reps, 1),
                      "sweeps_ms": {nm: round(sum(v[len(v) // 2:]) / len(v[len(v) // 2:]), 1) for nm, v in per.items()},
                      "mhz_median": sorted(clk[half:])[len(clk[half:]) // 2], "mhz_min": min(clk[half:]), "mhz_max": max(clk[half:]),
                      "power_median": round(sorted(pw[half:])[len(pw[half:]) // 2]), "power_max": round(max(pw)),
                      "temp_max": max(tmp), "reasons": reasons, "samples": len(clk)}), flush=True)
