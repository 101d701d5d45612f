"""Dev probe: streamed volume / failed guesses / time of the one-pass evaluation as a function of the extrapolation
safety factor ONE_PASS_GAMMA, against the two-sweep evaluation (bit-identical ranks asserted)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from snag_b200 import evaluate, ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_1m"
gammas = [float(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.5, 1.0, 1.5, 2.0, 3.0]
n, d, k, sigma, _ = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
emb, left, right = bench.synth_tables(n, d, sigma, dev)
X, xn = ops.prep_bf16(emb, left, True)
Y, yn = ops.prep_bf16(emb, right, True)
del emb


def timed(**kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = []
    ops.SWEEP_EVENT_SINK = ev
    e0.record()
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, True, **kw)
    e1.record()
    torch.cuda.synchronize()
    ops.SWEEP_EVENT_SINK = None
    return res, e0.elapsed_time(e1), {nm: round(a.elapsed_time(b), 2) for nm, a, b, *_ in ev}


samples = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [evaluate.SAMPLE_MAX]
ref, ms, ev = timed(one_pass=False)
ref, ms, ev = timed(one_pass=False)
print(json.dumps({"mode": "two_sweep", "ms": ms, "sweeps": ev, "rank_sweep": {k_: v for k_, v in ref.info["rank_sweep"].items()}}), flush=True)
for m_max in samples:
    evaluate.SAMPLE_MAX = m_max
    evaluate._ONE_PASS_CAP.clear()
    for gm in gammas:
        evaluate.ONE_PASS_GAMMA = gm
        res, ms, ev = timed(one_pass=True)
        res, ms, ev = timed(one_pass=True)
        same = bool(torch.equal(res.rank_l2r, ref.rank_l2r) and torch.equal(res.rank_r2l, ref.rank_r2l))
        print(json.dumps({"mode": "one_pass", "sample_max": m_max, "gamma": gm, "ms": ms, "same": same, "sweeps": ev,
                          "info": res.info["one_pass"]}), flush=True)
