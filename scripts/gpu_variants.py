"""A/B measurement of kernel build variants under SUSTAINED load on one GPU (power-capped clocks matter):
each variant runs in its own process (SNAG_B200_LIB), ~5 s of back-to-back sweeps per kernel, NVML sampled meanwhile."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import torch, pynvml
    from snag_b200 import evaluate, ops
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n, d, k = 300000, 1200, 10
    g = torch.Generator(device="cuda").manual_seed(1)
    centres = torch.randn((64, d), generator=g, device="cuda")
    X = torch.empty((n, 1216), dtype=torch.bfloat16, device="cuda"); Y = torch.empty_like(X)
    xn = torch.empty(n, device="cuda"); yn = torch.empty(n, device="cuda")
    for r0 in range(0, n, 50000):
        x = torch.randn((50000, d), generator=g, device="cuda") + centres[torch.randint(0, 64, (50000,), generator=g, device="cuda")]
        y = x + 7.0 * torch.randn((50000, d), generator=g, device="cuda")
        _, a = ops.prep_bf16(x, None, True, out=X[r0:r0 + 50000]); xn[r0:r0 + 50000] = a
        _, b = ops.prep_bf16(y, None, True, out=Y[r0:r0 + 50000]); yn[r0:r0 + 50000] = b
    nv = torch.zeros(n, device="cuda") + 0.5
    gg = torch.zeros(n, device="cuda") + 0.9
    cr = torch.zeros(n, dtype=torch.int32, device="cuda"); cc = torch.zeros(n, dtype=torch.int32, device="cuda")
    out = {}
    # two-sweep inputs: column thresholds from a sample pre-pass, exactly as evaluate._align_ranks_steps builds them
    m, cap = evaluate.two_sweep_plan(n, k)
    sel = torch.randperm(n, generator=torch.Generator(device="cpu").manual_seed(3408))[:m].sort()[0].cuda()
    part_s = ops.eval_rowtopk(Y, X.index_select(0, sel), yn, xn.index_select(0, sel), n, m)
    _, cand_s = ops.topk_merge_mean(part_s, k, want_nv=False, want_cand=True)
    colthr, colb = ops.col_threshold(cand_s, k, yn)
    scnt = ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, cap)[-1]
    out["cands_per_col"] = round(float(scnt.sum().item()) / n, 1)
    only = os.environ.get("SNAG_VARIANT_KERNELS", "mainloop,topk,rowcol,rank").split(",")
    for name, fn in (("mainloop", lambda: ops.sim_mainloop_only(X, Y, n, n)),
                     ("topk", lambda: ops.eval_rowtopk(X, Y, xn, yn, n, n)),
                     ("rowcol", lambda: ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, cap)),
                     ("rank", lambda: ops.eval_rank(X, Y, xn, yn, nv, nv, gg, gg, 0, 0, n, n, True, cr, cc))):
        if name not in only:
            continue
        fn(); torch.cuda.synchronize()
        clk, pw, stop = [], [], threading.Event()

        def sample():
            while not stop.is_set():
                clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000)
                stop.wait(0.05)
        th = threading.Thread(target=sample); th.start()
        reps = 30
        ev = [torch.cuda.Event(True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn(); ev[i + 1].record()
        torch.cuda.synchronize(); stop.set(); th.join()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        tail = ms[10:]
        out[name] = dict(ms_first=round(ms[0], 1), ms_tail=round(sum(tail) / len(tail), 1), tf_tail=round(2.0 * n * n * 1200 / (sum(tail) / len(tail)) / 1e9),
                         mhz_median=sorted(clk)[len(clk) // 2], mhz_min=min(clk), power_median=round(sorted(pw)[len(pw) // 2]))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(); sys.exit(0)
    vdir = os.path.join(ROOT, "snag_b200", "_variants")
    order = sys.argv[1:] or ["base", "wg4", "base"]
    for v in order:
        env = dict(os.environ, SNAG_B200_LIB=os.path.join(vdir, f"lib_{v}.so"))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, capture_output=True, text=True, timeout=300)
        print(v, (r.stdout.strip().splitlines() or ["<no output>"])[-1], r.stderr.strip()[-300:] if r.returncode else "", flush=True)
