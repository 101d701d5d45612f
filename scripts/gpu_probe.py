"""Which resource does an epilogue take away from the tensor pipe? Mainloop + TMEM read-out + synthetic per-strip load
(broadcast shared loads / dependent FMAs / shared stores), with the UMMA issuer's cycle counters. Development probe."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snag_b200 import ops
from snag_b200._lib import call, ptr, current_stream

sink = torch.zeros(1024, dtype=torch.int32, device="cuda")
dbg = torch.zeros(2 * 148 * 4, dtype=torch.int64, device="cuda")


def run(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    dbg.zero_()
    call("snag_debug_counters", ptr(dbg))
    fn(); torch.cuda.synchronize()
    call("snag_debug_counters", None)
    d = dbg.view(296, 4).double()
    tiles = d[:148, 3].mean().item()
    return dict(ms=round(ms, 3), cyc_per_tile=round(d[:148, 0].mean().item() / tiles), wait_acc=round(d[:148, 1].mean().item() / tiles),
                epi_proc=round(d[:148, 2].mean().item() / tiles / 16), epi_bar=round(d[148:, 0].mean().item() / tiles / 16),
                ghz=round(d[:148, 0].mean().item() / ms / 1e6, 3))


n, d = 100000, 1200
g = torch.Generator(device="cuda").manual_seed(1)
X, xn = ops.prep_bf16(torch.randn((n, d), generator=g, device="cuda"), None, True)
Y, yn = ops.prep_bf16(torch.randn((n, d), generator=g, device="cuda"), None, True)
dp = X.shape[1]
print(json.dumps(dict(kernel="mainloop", **run(lambda: ops.sim_mainloop_only(X, Y, n, n)))), flush=True)
for (l, a, s) in [(0, 0, 0), (0, -256, 0), (0, -1024, 0)]:
    r = run(lambda: call("snag_sim_readout_only", ptr(X), ptr(Y), n, n, dp, ptr(sink), l, a, s, current_stream()))
    print(json.dumps(dict(kernel="readout", lds=l, alu=a, sts=s, **r)), flush=True)
print(json.dumps(dict(kernel="topk", **run(lambda: ops.eval_rowtopk(X, Y, xn, yn, n, n)))), flush=True)
from snag_b200 import evaluate
res = evaluate.align_ranks(X, Y, xn, yn, n, 10, True)
cr = torch.zeros(n, dtype=torch.int32, device="cuda"); cc = torch.zeros(n, dtype=torch.int32, device="cuda")
print(json.dumps(dict(kernel="rank", **run(lambda: ops.eval_rank(X, Y, xn, yn, res.nv1, res.nv2, res.g, res.g, 0, 0, n, n, True, cr, cc)))), flush=True)

# fused row+column top-k sweep with different sample sizes (admission thresholds from m sampled rows)
for m in (8192, 32768):
    sel = torch.randperm(n, generator=torch.Generator(device="cpu").manual_seed(3408))[:m].sort()[0].cuda()
    part_s = ops.eval_rowtopk(Y, X.index_select(0, sel), yn, xn.index_select(0, sel), n, m)
    _, cand_s = ops.topk_merge_mean(part_s, 10, want_nv=False, want_cand=True)
    colthr, colb = ops.col_threshold(cand_s, 10, yn)
    cap = int(2.0 * 10 * n / m * n / 148) + 4096
    print(json.dumps(dict(kernel="fused", m=m, **run(lambda: ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, cap)))), flush=True)
