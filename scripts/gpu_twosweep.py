"""Development check of the two-sweep CSLS path against the three-sweep path (bitwise) + timings."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snag_b200 import evaluate, ops


def clustered(n, d, sigma, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    centres = torch.randn((64, d), generator=g, device="cuda")
    x = torch.randn((n, d), generator=g, device="cuda") + centres[torch.randint(0, 64, (n,), generator=g, device="cuda")]
    y = x + sigma * torch.randn((n, d), generator=g, device="cuda")
    return x, y


def t(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (n, d, k) in [(70000, 300, 10), (70000, 1800, 10), (100000, 1200, 10), (100000, 1200, 3), (250000, 1200, 10)]:
    x, y = clustered(n, d, 8.0, 5)
    X, xn = ops.prep_bf16(x, None, True)
    Y, yn = ops.prep_bf16(y, None, True)
    del x, y
    a = evaluate.align_ranks(X, Y, xn, yn, n, k, True, two_sweep=False)
    b = evaluate.align_ranks(X, Y, xn, yn, n, k, True, two_sweep=True)
    m, cap = evaluate.two_sweep_plan(n, k)
    # candidate statistics of the fused sweep
    gsel = torch.Generator(device="cpu").manual_seed(3408)
    sel = torch.randperm(n, generator=gsel)[:m].sort()[0].cuda()
    part_s = ops.eval_rowtopk(Y, X.index_select(0, sel), yn, xn.index_select(0, sel), n, m)
    _, cand_s = ops.topk_merge_mean(part_s, k, want_nv=False, want_cand=True)
    colthr, colb = ops.col_threshold(cand_s, k, yn)
    part, pidx, stream, srow, scnt = ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, cap)
    cval, cidx, ovf = ops.col_cand_reduce(stream, srow, scnt, n, k)
    cnt = torch.full((1,), float(scnt.sum().item()) / n)
    ms_reduce = t(lambda: ops.col_cand_reduce(stream, srow, scnt, n, k))
    ms_pre = t(lambda: ops.eval_rowtopk(Y, X.index_select(0, sel), yn, xn.index_select(0, sel), n, m))
    ms_main = t(lambda: ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, cap))
    ms_old = t(lambda: ops.eval_rowtopk(X, Y, xn, yn, n, n))
    ms3 = t(lambda: evaluate.align_ranks(X, Y, xn, yn, n, k, True, two_sweep=False))
    ms2 = t(lambda: evaluate.align_ranks(X, Y, xn, yn, n, k, True, two_sweep=True))
    print(json.dumps(dict(n=n, d=d, k=k, m=m, cap=cap, launches=(a.launches, b.launches),
                          nv1_equal=bool(torch.equal(a.nv1, b.nv1)), nv2_equal=bool(torch.equal(a.nv2, b.nv2)),
                          nv2_maxdiff=float((a.nv2 - b.nv2).abs().max()),
                          l2r_equal=bool(torch.equal(a.rank_l2r, b.rank_l2r)), r2l_equal=bool(torch.equal(a.rank_r2l, b.rank_r2l)),
                          cand_mean=float(cnt.float().mean()), cand_max=int(cnt.max()), expect=k * n / m, overflow=int(ovf.item()),
                          stream_max=int(scnt.max()), stream_min=int(scnt.min()), ms_reduce=ms_reduce,
                          ms_prepass=ms_pre, ms_fused_sweep=ms_main, ms_plain_rowtopk=ms_old, ms_eval3=ms3, ms_eval2=ms2,
                          speedup=ms3 / ms2)), flush=True)
    del X, Y, part, pidx, stream, srow
    torch.cuda.empty_cache()
