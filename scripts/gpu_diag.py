"""First-contact diagnostics for the CUDA path on a real B200. Each stage runs in its own process under a
timeout so that a trap or a hang in one kernel does not take the others down. Results go to stdout and
gpurun_out/diag.jsonl. This is a development tool, not part of the test-suite (tests/ holds the parity tests).

    python scripts/gpu_diag.py [stage ...]
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def emit(stage, **kw):
    rec = {"stage": stage, **kw}
    print(json.dumps(rec), flush=True)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")


# ------------------------------------------------------------------------------------------------ stages
def stage_device():
    import torch
    from snag_b200 import _lib
    lib = _lib.load()
    emit("device", name=torch.cuda.get_device_name(0), cc=torch.cuda.get_device_capability(0),
         check=lib.snag_device_check(), sms=lib.snag_num_sms())


def _rand_operand(n, d, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)


def stage_prep():
    import torch
    from snag_b200 import ops
    for (n, d) in [(5, 64), (1000, 300), (777, 1200), (513, 1800)]:
        e = _rand_operand(n, d, 1)
        X, xn = ops.prep_bf16(e, None, True)
        ref = torch.nn.functional.normalize(e).to(torch.bfloat16)
        diff = (X[:, :d].float() - ref.float()).abs().max().item()
        pad = X[:, d:].float().abs().max().item() if X.shape[1] > d else 0.0
        xn_ref = (X.double() ** 2).sum(1).float()
        emit("prep", n=n, d=d, dpad=X.shape[1], max_diff=diff, pad_max=pad,
             xn_max_diff=(xn - xn_ref).abs().max().item())


def stage_gemm():
    """sim_write mode 0 against torch fp32 matmul on the same bf16 values."""
    import torch
    from snag_b200 import ops
    shapes = [(128, 256, 64), (128, 256, 128), (128, 256, 320), (256, 512, 64), (300, 500, 300), (1000, 1300, 1200),
              (4100, 4200, 1800)]
    for (n1, n2, d) in shapes:
        x = _rand_operand(n1, d, 2)
        y = _rand_operand(n2, d, 3)
        X, xn = ops.prep_bf16(x, None, True)
        Y, yn = ops.prep_bf16(y, None, True)
        S = ops.sim_write(X, Y, None, None, n1, n2, 0)
        torch.cuda.synchronize()
        ref = X.float() @ Y.float().t()
        err = (S - ref).abs()
        rec = dict(n1=n1, n2=n2, d=d, max_err=err.max().item(), mean_err=err.mean().item(),
                   ref_absmax=ref.abs().max().item())
        if rec["max_err"] > 1e-3:
            bad = (err > 1e-3)
            rec["bad_frac"] = bad.float().mean().item()
            rec["bad_rows"] = bad.any(1).nonzero().flatten()[:16].tolist()
            rec["bad_cols"] = bad.any(0).nonzero().flatten()[:16].tolist()
            rec["sample_S"] = S[:2, :8].tolist()
            rec["sample_ref"] = ref[:2, :8].tolist()
        emit("gemm", **rec)
        D1 = ops.sim_write(X, Y, xn, yn, n1, n2, 1)
        dref = torch.clamp((xn[:, None] + yn[None, :]) - 2.0 * S, min=0.0)
        emit("dist", n1=n1, n2=n2, d=d, max_err=(D1 - dref).abs().max().item())


def _torch_eval_reference(X, Y, n, k, use_csls=True):
    """Reference chain in torch fp32 on the GPU (same bf16 values), ranks by counting with the stable tie-break."""
    import torch
    Xf, Yf = X[:n].float(), Y[:n].float()
    xn = (Xf.double() ** 2).sum(1).float()
    yn = (Yf.double() ** 2).sum(1).float()
    s = (Xf.double() @ Yf.double().t()).float()
    d = torch.clamp((xn[:, None] + yn[None, :]) - 2.0 * s, min=0.0)
    if use_csls:
        c = 1 - d
        t1 = torch.topk(c, k, dim=1)[0]
        t2 = torch.topk(c.t(), k, dim=1)[0]
        nv1 = t1[:, 0].clone()
        nv2 = t2[:, 0].clone()
        for t in range(1, k):
            nv1 = nv1 + t1[:, t]
            nv2 = nv2 + t2[:, t]
        nv1, nv2 = nv1 / k, nv2 / k
        dist = 1 - ((2 * c - nv1[:, None]) - nv2[None, :])
    else:
        nv1 = nv2 = None
        dist = d
    g = dist.diagonal()
    idx = torch.arange(n, device=X.device)
    lt_row = (dist < g[:, None]) | ((dist == g[:, None]) & (idx[None, :] < idx[:, None]))
    lt_col = (dist < g[None, :]) | ((dist == g[None, :]) & (idx[:, None] < idx[None, :]))
    lt_row[idx, idx] = False
    lt_col[idx, idx] = False
    return dict(nv1=nv1, nv2=nv2, g=g, rank_l2r=lt_row.sum(1).int(), rank_r2l=lt_col.sum(0).int(), dist=dist)


def _clustered(n, d, sigma, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    centres = torch.randn((64, d), generator=g, device="cuda")
    assign = torch.randint(0, 64, (n,), generator=g, device="cuda")
    x = torch.randn((n, d), generator=g, device="cuda") + centres[assign]
    y = x + sigma * torch.randn((n, d), generator=g, device="cuda")
    return x, y


def stage_eval():
    import torch
    from snag_b200 import evaluate, ops
    for (n, d, k, csls) in [(300, 128, 3, True), (1000, 300, 10, True), (2500, 1200, 10, True), (2500, 1200, 10, False),
                            (5000, 1800, 3, True)]:
        x, y = _clustered(n, d, 8.0, 5)
        X, xn = ops.prep_bf16(x, None, True)
        Y, yn = ops.prep_bf16(y, None, True)
        res = evaluate.align_ranks(X, Y, xn, yn, n, k, csls, want_top3=True)
        torch.cuda.synchronize()
        ref = _torch_eval_reference(X, Y, n, k, csls)
        rec = dict(n=n, d=d, k=k, csls=csls)
        if csls:
            rec["nv1_err"] = (res.nv1 - ref["nv1"]).abs().max().item()
            rec["nv2_err"] = (res.nv2 - ref["nv2"]).abs().max().item()
        rec["g_err"] = (res.g - ref["g"]).abs().max().item()
        rec["l2r_mismatch"] = int((res.rank_l2r != ref["rank_l2r"]).sum().item())
        rec["r2l_mismatch"] = int((res.rank_r2l != ref["rank_r2l"]).sum().item())
        rec["l2r_maxabs"] = int((res.rank_l2r - ref["rank_l2r"]).abs().max().item())
        t3 = torch.topk(ref["dist"], 3, dim=1, largest=False)[1].int()
        rec["top3_mismatch"] = int((res.top3_idx != t3).any(1).sum().item())
        m = evaluate.metrics_from_ranks(res.rank_l2r)
        rec["hits1"] = float(m.acc[0])
        rec["mrr"] = m.mrr
        emit("eval", **rec)


def stage_icl():
    import torch
    from snag_b200 import ops
    for (B, d, tau) in [(100, 64, 0.1), (1000, 300, 0.1), (3500, 300, 0.1), (1500, 1200, 0.1)]:
        a = torch.nn.functional.normalize(_rand_operand(B, d, 7))
        b = torch.nn.functional.normalize(a + 0.5 * torch.nn.functional.normalize(_rand_operand(B, d, 8)))
        Bp = ops.round_up(B, 256)
        A, _ = ops.prep_bf16(a, None, False, rows_pad_to=256)
        Bm, _ = ops.prep_bf16(b, None, False, rows_pad_to=256)
        Ya = torch.cat([Bm, A], 0).contiguous()
        lse, nll, pos = ops.icl_side(A, Ya, B, Bp, 1.0 / tau)
        torch.cuda.synchronize()
        Af, Bf = A[:B].float(), Bm[:B].float()
        lab = torch.cat([Af @ Bf.t(), Af @ Af.t() - 1e9 * torch.eye(B, device="cuda")], 1) / tau
        lse_ref = torch.logsumexp(lab, 1)
        nll_ref = lse_ref - lab.diagonal()
        emit("icl", B=B, d=d, lse_err=(lse - lse_ref).abs().max().item(), nll_err=(nll - nll_ref).abs().max().item(),
             nll_mean=nll_ref.mean().item())


def stage_noise():
    import torch
    from snag_b200 import ops
    N, F = 5000, 1000
    x = _rand_operand(N, F, 11)
    mean, std = ops.col_mean_std(x)
    emit("colstats", mean_err=(mean - x.mean(0)).abs().max().item(), std_err=(std - x.std(0)).abs().max().item())
    valid = (torch.rand(N, device="cuda") < 0.8).to(torch.uint8)
    m2, s2 = ops.col_mean_std(x, valid)
    xs = x[valid.bool()]
    emit("colstats_subset", mean_err=(m2 - xs.mean(0)).abs().max().item(), std_err=(s2 - xs.std(0)).abs().max().item())
    mask = (torch.rand(N, device="cuda") < 0.2)
    z = torch.randn((int(mask.sum()), F), device="cuda")
    out = ops.noise_mask(x, mean, std, 0.2, 0.7, mask=mask.to(torch.uint8), zsel=z)
    ref = x.clone()
    ref[mask] = (1.0 - 0.7) * x[mask] + 0.7 * (mean + std * z)
    emit("noise_injected", max_err=(out - ref).abs().max().item(), exact=bool((out == ref).all().item()))
    out2 = ops.noise_mask(x, mean, std, 0.2, 0.7, seed=3408)
    changed = (out2 != x).any(1)
    zhat = ((out2[changed] - 0.3 * x[changed]) / 0.7 - mean) / std
    emit("noise_philox", frac=changed.float().mean().item(), z_mean=zhat.mean().item(), z_std=zhat.std().item(),
         untouched_exact=bool((out2[~changed] == x[~changed]).all().item()))
    en = ops.gauss_fill(mean, std, N, 3408)
    zz = (en - mean) / std
    emit("gauss_fill", z_mean=zz.mean().item(), z_std=zz.std().item(), z_kurt=(zz ** 4).mean().item())
    e = _rand_operand(N, 300, 12)
    nz = _rand_operand(N, 300, 13)
    mk = ops.philox_rowmask(N, 0.1, 3408, "cuda")
    o = ops.rowblend_fwd(e, nz, mk, 1.0 - 0.35, 0.35)
    r = e.clone()
    r[mk.bool()] = (1.0 - 0.7 * 0.5) * e[mk.bool()] + 0.7 * 0.5 * nz[mk.bool()]
    gi = ops.rowblend_bwd(e, mk, 0.65)
    rg = e.clone()
    rg[mk.bool()] = 0.65 * e[mk.bool()]
    emit("rowblend", frac=mk.float().mean().item(), fwd_exact=bool((o == r).all().item()),
         bwd_exact=bool((gi == rg).all().item()))


def stage_perf():
    """First timing numbers for the sweeps (CUDA events, after warm-up)."""
    import torch
    from snag_b200 import evaluate, ops
    for (n, d, k) in [(10500, 1200, 10), (10500, 1800, 10), (32768, 300, 10), (32768, 1200, 10), (100000, 1200, 10)]:
        x, y = _clustered(n, d, 8.0, 5)
        X, xn = ops.prep_bf16(x, None, True)
        Y, yn = ops.prep_bf16(y, None, True)
        del x, y
        dpad = X.shape[1]

        def t(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_null = t(lambda: ops.sim_mainloop_only(X, Y, n, n))
        emit("perf_mainloop", n=n, d=d, dpad=dpad, ms=ms_null, tf=2.0 * n * n * dpad / ms_null / 1e9)
        ms_topk = t(lambda: ops.eval_rowtopk(X, Y, xn, yn, n, n))
        res = evaluate.align_ranks(X, Y, xn, yn, n, k)
        cr = torch.zeros(n, dtype=torch.int32, device="cuda")
        cc = torch.zeros(n, dtype=torch.int32, device="cuda")
        ms_rank = t(lambda: ops.eval_rank(X, Y, xn, yn, res.nv1, res.nv2, res.g, res.g, 0, 0, n, n, True, cr, cc))
        ms_all = t(lambda: evaluate.align_ranks(X, Y, xn, yn, n, k))
        fl = 2.0 * n * n * dpad
        emit("perf", n=n, d=d, dpad=dpad, plan=ops.sim_plan(n, n, dpad), ms_topk=ms_topk, tf_topk=fl / ms_topk / 1e9,
             ms_rank=ms_rank, tf_rank=fl / ms_rank / 1e9, ms_eval=ms_all, pairs_per_s=n * n / ms_all * 1e3,
             hits1=float(evaluate.metrics_from_ranks(res.rank_l2r).acc[0]))


STAGES = {"device": stage_device, "prep": stage_prep, "gemm": stage_gemm, "eval": stage_eval, "icl": stage_icl,
          "noise": stage_noise, "perf": stage_perf}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    todo = sys.argv[1:] or list(STAGES)
    for s in todo:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=240,
                               capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                emit(s, error=f"exit {r.returncode}", stderr=r.stderr[-1500:])
        except subprocess.TimeoutExpired as e:
            emit(s, error="timeout", stdout=(e.stdout or b"")[-500:].decode() if isinstance(e.stdout, bytes) else str(e.stdout)[-500:])
        emit(s + "_done", seconds=round(time.time() - t0, 1))
