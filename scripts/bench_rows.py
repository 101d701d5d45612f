"""Per-row measurements of SURVEY 8(a) outside the two tensor-core sweeps bench.py reports: the bandwidth kernels
(noise mask, column statistics, entity blend, bf16 prologue), the materialising drop-ins (pairwise_distances,
csls_sim) and the small evaluation glue — each against the HBM roofline (MEASURED_PEAKS.json hbm_gbs, measured copy
bandwidth) with the reference's own torch op sequence timed on the host cores beside it.

    python scripts/bench_rows.py [--shape c1|c3] > gpurun_out/rows.jsonl

One JSON line per row: algorithmic bytes per launch (SURVEY 8(d) figure), CUDA-event time per launch on the launching
stream (median of `reps`, L2 flushed between launches by writing a 512 MB buffer), achieved GB/s, fraction of peak,
CPU seconds of the reference op sequence (torch CPU, all host threads) on the same shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from snag_b200 import evaluate, fusion, loss, mining, noise, ops, seeds


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "B200_PROFILING.md fallback"


def gpu_time(fn, reps=9):
    flush = torch.empty((512 << 20,), dtype=torch.uint8, device="cuda")
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def cpu_time(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c1", choices=["c1", "c3"])
    args = ap.parse_args()
    N, F_img, n_test = (39594, 2048, 10500) if args.shape == "c1" else (27793, 4096, 10277)
    D = 1200
    hbm, src = peaks()
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator(device="cuda").manual_seed(3408)
    rows = []

    tf_peak = 1660.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        tf_peak = float(json.load(open(pk)).get("bf16_tflops", tf_peak))

    def emit(row, kernel, nbytes, ms, cpu_s, note="", flops=None):
        """HBM-bound rows: algorithmic bytes against the measured copy bandwidth; tensor-bound rows (flops given):
        algorithmic flops against the measured burst bf16 peak (launches of a few ms timed alone)."""
        rec = {"row": row, "kernel": kernel, "shape": args.shape, "ms": round(ms, 4)}
        if flops is None:
            rec.update({"bound": "hbm", "algorithmic_bytes": int(nbytes), "achieved_gbs": round(nbytes / ms / 1e6, 1),
                        "peak_gbs": hbm, "frac": round(nbytes / ms / 1e6 / hbm, 3), "peak_source": src})
        else:
            rec.update({"bound": "tensor", "algorithmic_flops": float(flops), "achieved_tflops": round(flops / ms / 1e9, 1),
                        "peak_tflops": tf_peak, "frac": round(flops / ms / 1e9 / tf_peak, 3),
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)"})
        rec.update({"cpu_s": None if cpu_s is None else round(cpu_s, 4), "cpu_threads": torch.get_num_threads(),
                    "speedup_vs_cpu": None if cpu_s is None else round(cpu_s * 1e3 / ms, 1), "note": note})
        rows.append(rec)
        print(json.dumps(rec), flush=True)

    # ------------------------------------------------------------------ a1 noise mask (model/SNAG.py:66-75)
    feats = {"rel": torch.poisson(torch.full((N, 1000), 0.05, device="cuda"), generator=g),
             "att": torch.bernoulli(torch.full((N, 1000), 0.01, device="cuda"), generator=g),
             "img": torch.nn.functional.normalize(torch.randn((N, F_img), generator=g, device="cuda"))}
    r, rho = 0.2, 0.7

    def ref_add_noise(x, mean, std):
        x = x.clone()
        m = torch.rand(x.shape[0]) < r
        sel = x[m]
        x[m] = (1 - rho) * sel + rho * (mean + std * torch.randn_like(sel))
        return x

    for name, x in feats.items():
        mean, std = ops.col_mean_std(x)
        out = torch.empty_like(x)
        ms = gpu_time(lambda: ops.noise_mask(x, mean, std, r, rho, out=out, seed=7))
        xc, mc, sc = x.cpu(), mean.cpu(), std.cpu()
        emit("a1", f"noise_mask_kernel[{name} {tuple(x.shape)}]", 8 * x.numel(), ms, cpu_time(lambda: ref_add_noise(xc, mc, sc)),
             "read x + write noisy copy, Philox drawn in-kernel")
        ms = gpu_time(lambda: ops.col_mean_std(x))
        emit("a3", f"col_stats[{name} {tuple(x.shape)}]", 4 * x.numel(), ms,
             cpu_time(lambda: (torch.mean(xc, dim=0), torch.std(xc, dim=0))), "one read, fp64 partial sums")

    # the same two kernels on a 10x taller matrix: launch latency amortised, the kernel's own bandwidth shows
    xb = torch.randn((400_000, 1000), generator=g, device="cuda")
    mb, sb = ops.col_mean_std(xb)
    ob = torch.empty_like(xb)
    ms = gpu_time(lambda: ops.noise_mask(xb, mb, sb, r, rho, out=ob, seed=7), reps=5)
    emit("a1", "noise_mask_kernel[400000 x 1000] (size sweep)", 8 * xb.numel(), ms, None, "3.2 GB of traffic per launch")
    ms = gpu_time(lambda: ops.col_mean_std(xb), reps=5)
    emit("a3", "col_stats[400000 x 1000] (size sweep)", 4 * xb.numel(), ms, None)
    del xb, ob

    # ------------------------------------------------------------------ a2 entity noise + a4 blend (SNAG.py:94-98, SNAG_tools.py:122-129)
    ent = torch.randn((N, 300), generator=g, device="cuda") / N ** 0.5
    em, es = ops.col_mean_std(ent)
    ms = gpu_time(lambda: ops.gauss_fill(em, es, N, 11))
    emc, esc, entc = em.cpu(), es.cpu(), ent.cpu()
    emit("a2", f"gauss_fill[{N}x300]", 4 * N * 300, ms, cpu_time(lambda: emc + esc * torch.randn_like(entc)), "write only")
    noise_t = ops.gauss_fill(em, es, N, 11)
    mask = ops.philox_rowmask(N, r * 0.5, 13, ent.device)
    a, c = float(np.float32(1 - rho * 0.5)), float(np.float32(rho * 0.5))
    ms = gpu_time(lambda: ops.rowblend_fwd(ent, noise_t, mask, a, c))
    nc, mcpu = noise_t.cpu(), mask.cpu().bool()

    def ref_blend():
        e = entc.clone()
        e[mcpu] = a * e[mcpu] + c * nc[mcpu]
        return e
    emit("a4", f"rowblend_fwd[{N}x300]", 8 * N * 300 + 4 * N * 300 * (r * 0.5), ms, cpu_time(ref_blend),
         "read e + write e'; noise rows read only where masked")
    gr = torch.randn_like(ent)
    ms = gpu_time(lambda: ops.rowblend_bwd(gr, mask, a))
    emit("a4", f"rowblend_bwd[{N}x300]", 8 * N * 300, ms, None)

    # ------------------------------------------------------------------ prologue: gather + normalise + bf16 + norm
    emb = torch.randn((N, D), generator=g, device="cuda")
    idx = torch.randperm(N, generator=g, device="cuda")[:n_test].contiguous()
    ms = gpu_time(lambda: ops.prep_bf16(emb, idx, True))
    ec, ic = emb.cpu(), idx.cpu()
    emit("prologue", f"prep_bf16[{n_test}x{D}]", n_test * (4 * D + 2 * ops.round_up(D, 64)), ms,
         cpu_time(lambda: torch.nn.functional.normalize(ec)[ic]), "main.py:379 normalises all N rows; only the gathered rows here")

    # ------------------------------------------------------------------ a8 / a9 materialising drop-ins
    x, y = emb[idx], emb[torch.randperm(N, generator=g, device="cuda")[:n_test]]
    x, y = torch.nn.functional.normalize(x), torch.nn.functional.normalize(y)
    ms = gpu_time(lambda: evaluate.pairwise_distances(x, y), reps=5)
    xc, yc = x.cpu(), y.cpu()

    def ref_pd():
        xn = (xc ** 2).sum(1).view(-1, 1)
        yn = (yc ** 2).sum(1).view(1, -1)
        return torch.clamp(xn + yn - 2.0 * torch.mm(xc, yc.t()), 0.0, np.inf)
    emit("a8", f"pairwise_distances[{n_test}^2, D={D}]", 4 * n_test * n_test, ms, cpu_time(ref_pd, 2),
         "HBM-write bound: the fp32 [n,n] output; includes both bf16 prologues")
    sim = 1 - evaluate.pairwise_distances(x, y)
    ms = gpu_time(lambda: evaluate.csls_sim(sim, 10), reps=5)
    sc = sim.cpu()

    def ref_csls():
        nv1 = torch.mean(torch.topk(sc, 10)[0], 1)
        nv2 = torch.mean(torch.topk(sc.t(), 10)[0], 1)
        return (2 * sc.t() - nv1).t() - nv2
    emit("a9", f"csls_sim[{n_test}^2, k=10] (materialised drop-in)", 12 * n_test * n_test, ms, cpu_time(ref_csls, 2),
         "2 reads + 1 write of the matrix is the algorithmic minimum; executed: 3 reads + 1 write")
    del sim, sc
    # ------------------------------------------------------------------ f2 fusion-output embeddings (SNAG_tools.py:44-49)
    M = 4
    embs = [torch.randn((N, 300), generator=g, device="cuda") for _ in range(M)]
    wn = torch.softmax(torch.randn((N, M), generator=g, device="cuda"), 1)
    wg = torch.softmax(torch.randn((6,), generator=g, device="cuda"), 0)
    ms = gpu_time(lambda: ops.joint_fuse_fwd(embs, wn, wg))
    ecs, wnc, wgc = [e.cpu() for e in embs], wn.cpu(), wg.cpu()

    def ref_joint():
        j = torch.cat([wnc[:, m].unsqueeze(1) * torch.nn.functional.normalize(ecs[m]) for m in range(M)], dim=1)
        jf = torch.cat([wgc[m] * torch.nn.functional.normalize(ecs[m]) for m in range(M)], dim=1)
        return j, jf
    emit("f2", f"joint_fuse_fwd[{N} x {M}x300 -> 2 x {M * 300}]", 4 * N * 300 * M * 3, ms, cpu_time(ref_joint),
         "read M tables once + write joint_emb and joint_emb_fz; the reference makes 4M+2 passes")
    dj = torch.randn((N, 300 * M), generator=g, device="cuda")
    djf = torch.randn((N, 300 * M), generator=g, device="cuda")
    ms = gpu_time(lambda: ops.joint_fuse_bwd(embs, wn, wg, dj, djf))
    emit("f2", f"joint_fuse_bwd[{N} x {M}x300]", 4 * N * 300 * M * 4, ms, None, "read e, dJ, dJfz + write de (each table read twice: norm pass + apply pass hits L2)")
    del embs, dj, djf

    # ------------------------------------------------------------------ f1 link mining (SNAG.py:192-208): fused mutual-NN sweep
    n_l = n_r = N // 2 - 2250 if args.shape == "c1" else N // 2 - 1285          # the non-train entities of each KG
    xl = torch.nn.functional.normalize(torch.randn((n_l, D), generator=g, device="cuda"))
    yr = torch.nn.functional.normalize(xl[:n_r] + 0.5 * torch.nn.functional.normalize(torch.randn((n_r, D), generator=g, device="cuda")))
    ms = gpu_time(lambda: mining.mutual_nearest(xl, yr, normalize=False), reps=5)
    xlc, yrc = xl.cpu(), yr.cpu()

    def ref_mine():
        xn = (xlc ** 2).sum(1).view(-1, 1)
        yn = (yrc ** 2).sum(1).view(1, -1)
        d = torch.clamp(xn + yn - 2.0 * torch.mm(xlc, yrc.t()), 0.0, np.inf)
        return torch.argmin(d, dim=1), torch.argmin(d.t(), dim=1)
    flops = 2.0 * n_l * n_r * D
    rec_ms = ms
    emit("f1", f"mutual_nearest[{n_l} x {n_r}, D={D}] (two top-k sweeps + canonical re-score of the candidates)", 0, rec_ms,
         cpu_time(ref_mine, 2), "algorithmic = one pass over the distance matrix (2*n_l*n_r*D); executed: two sweeps", flops=flops)
    # ------------------------------------------------------------------ f4 unsupervised seeds (src/data.py:367-402): global top-K
    Kq = 100_000                                                             # unsup_k = 1000 (config.py:70) x 100
    ms = gpu_time(lambda: seeds.topk_similarity_entries(xl, yr, Kq), reps=5)

    def ref_topk():
        sim = xlc.mm(yrc.t())
        vals, ind = sim.view(-1).topk(Kq)
        return ind // sim.shape[1], ind % sim.shape[1]
    emit("f4", f"topk_similarity_entries[{n_l} x {n_r}, D={D}, K={Kq}] (pool sweep + thresholded sweep + re-score + sort)",
         0, ms, cpu_time(ref_topk, 2), "algorithmic = one pass over the similarity matrix; executed: two sweeps", flops=flops)
    del xl, yr
    # ------------------------------------------------------------------ a5 / a6 one loss call fwd+bwd (SNAG_loss.py:58-128, 148-202)
    B = 3500
    links = torch.stack([torch.randperm(N // 2, generator=g, device="cuda")[:B],
                         N // 2 + torch.randperm(N // 2, generator=g, device="cuda")[:B]], 1)
    src_emb = torch.randn((N, 300), generator=g, device="cuda", requires_grad=True)
    tar_emb = torch.randn((N, D), generator=g, device="cuda")
    icl = loss.icl_loss(tau=0.1, ab_weight=0.5, n_view=2)
    ial = loss.ial_loss(tau=4.0, ab_weight=0.5, zoom=0.1, reduction="mean")

    def run_icl():
        src_emb.grad = None
        icl(src_emb, links).backward()

    def run_ial():
        src_emb.grad = None
        ial(src_emb, tar_emb, links).backward()
    ms = gpu_time(run_icl, reps=5)
    emit("a5", f"icl_loss fwd+bwd[B={B}, D=300] (one call, eager)", 0, ms, None,
         "algorithmic: 6*B^2*D fwd + 14*B^2*D bwd (SURVEY 8d); includes the gather/normalise/scatter glue and launch gaps",
         flops=20.0 * B * B * 300)
    ms = gpu_time(run_ial, reps=5)
    emit("a6", f"ial_loss fwd+bwd[B={B}, src D=300, tar D={D}] (one call, eager)", 0, ms, None,
         "fused row-wise KL: 4 log-sum-exp sweeps + 2 softmax writers + 4 split-K products fwd, 4 dL/dlogits sweeps + 2 products bwd; algorithmic = 12*B^2*Dbar fwd (SURVEY 8d) x 3",
         flops=12.0 * B * B * (300 + D) / 2 * 3)
    out = os.path.join(ROOT, "gpurun_out", f"rows_{args.shape}.jsonl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        for rec in rows:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
