#!/usr/bin/env python
"""bench.py — alignment-evaluation throughput (similarity + CSLS + rank -> Hits@k / MR / MRR) on synthetic
DBP15K/FBDB15K-shaped data, the headline metric of BASELINE.json, on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl snag|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one full evaluation of n aligned entity pairs: gather + L2-normalise + bf16 cast of both embedding
tables, CSLS row/column neighbourhood sweeps, ground-truth scores, rank-count sweep (both directions). With N > 1
the targets are sharded over the ranks (strong scaling: the workload is fixed) and the per-row candidates /
counters are exchanged with NCCL. Rank 0 prints ONE JSON line (contract in the task statement):
  value      whole-job pairs/s, embeddings resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e        the same metric through the public API from pinned HOST buffers: H2D of the embeddings, evaluation,
             D2H of the ranks, host Hits/MR/MRR reduction — all inside the timed region
  roofline   the dominant kernel (fused tcgen05 sweep) against the measured bf16 tensor peak
  cpu_baseline  the oracle port timed on this host's cores on a bounded sample of the same workload
`--impl reference` times the CPU port alone (rank 0 only) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "align-eval entity pairs/sec (sim+CSLS+rank)"
UNIT = "pairs/s"

# name -> (n pairs, per-modality joint width D, csls k, noise sigma of the synthetic targets, description)
WORKLOADS = {
    "c4_1m": (1_000_000, 1200, 10, 6.0, "configs[3] alignment evaluation 1M x 1M entities, D=1200, CSLS k=10"),
    "c4_500k": (500_000, 1200, 10, 6.5, "configs[3] alignment evaluation 500k x 500k, D=1200, CSLS k=10"),
    "c4_200k": (200_000, 1200, 10, 7.0, "configs[3] alignment evaluation 200k x 200k, D=1200, CSLS k=10"),
    "c4_100k": (100_000, 1200, 10, 8.0, "configs[3] alignment evaluation 100k x 100k, D=1200, CSLS k=10"),
    "c1": (10_500, 1200, 10, 8.0, "configs[0] DBP15K ja_en-shaped eval, 10 500 test pairs, D=1200, CSLS k=10"),
    "c2": (10_500, 1800, 10, 8.0, "configs[1] DBP15K fr_en-shaped + surface eval, 10 500 test pairs, D=1800, CSLS k=10"),
    "c3": (10_277, 1200, 10, 8.0, "configs[2] FBDB15K-shaped eval, 10 277 test pairs, D=1200, CSLS k=10"),
}
# loss-layer slice of one training step: name -> (batch B, modalities M, width per modality, description)
TRAIN_WORKLOADS = {
    "c5_train": (16384, 6, 300, "configs[4] large-batch ICL: B=16384 in-batch negatives, 6 streams x 300 (+ joint 1800), "
                                "2 + 2*6 icl_loss calls fwd+bwd per step"),
    "c2_train": (3500, 6, 300, "configs[1] DBP15K fr_en + surface: B=3500, 6 streams x 300 (+ joint 1800), 14 icl_loss calls"),
    "c1_train": (3500, 4, 300, "configs[0] DBP15K ja_en: B=3500, 4 streams x 300 (+ joint 1200), 10 icl_loss calls"),
}
TRAIN_METRIC = "train steps/sec (loss-layer slice: 2+2M icl_loss fwd+bwd)"
DEFAULT_WORKLOAD = "c4_1m"
SEED = 3408                       # the reference's scripted seed (run.sh:2)
CPU_SAMPLE_N = 4096               # bounded sample for the CPU legs: a CPU_SAMPLE_N x CPU_SAMPLE_N sub-problem


# ================================================================================================ helpers
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tensor_sustained": float(p["bf16_tflops_sustained"]), "tensor_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured)"}
    return {"tensor_sustained": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "B200_PROFILING.md (fallback)"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, 100 ms period)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.index, self.samples, self.mask, self.max_mhz = index, [], 0, None
        self.power_w, self.temp_c = [], []
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:          # NVML unavailable: report nulls rather than fail the benchmark
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power_w.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.temp_c.append(int(self.nv.nvmlDeviceGetTemperature(self.h, self.nv.NVML_TEMPERATURE_GPU)))
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        pw, tc = sorted(self.power_w), sorted(self.temp_c)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": [name for bit, name in self.REASONS.items() if self.mask & bit], "samples": len(s),
                "sm_mhz_min": (s[0] if s else None), "power_w_median": (pw[len(pw) // 2] if pw else None),
                "temp_c_max": (tc[-1] if tc else None)}


def synth_tables(n: int, d: int, sigma: float, device, chunk: int = 65536):
    """SURVEY 8(d) generator: 64 cluster centres c ~ N(0, I); x_i = N(0, I) + c_g(i); y_i = x_i + sigma N(0, I).
    Returns the joint embedding table final_emb [2n, d] fp32 (sources then targets) and the two index vectors."""
    import torch
    g = torch.Generator(device=device).manual_seed(SEED)
    centres = torch.randn((64, d), generator=g, device=device)
    emb = torch.empty((2 * n, d), dtype=torch.float32, device=device)
    for r0 in range(0, n, chunk):
        r1 = min(r0 + chunk, n)
        x = torch.randn((r1 - r0, d), generator=g, device=device)
        x += centres[torch.randint(0, 64, (r1 - r0,), generator=g, device=device)]
        emb[r0:r1] = x
        emb[n + r0:n + r1] = x + sigma * torch.randn((r1 - r0, d), generator=g, device=device)
    left = torch.arange(0, n, device=device, dtype=torch.int64)
    right = torch.arange(n, 2 * n, device=device, dtype=torch.int64)
    return emb, left, right


# ================================================================================================ CPU legs
def cpu_port_sample(n_s: int, d: int, k: int, sigma: float, steps: int, warmup: int):
    """Times the oracle port (oracle/snag_oracle.c, all host threads) on an n_s x n_s sub-problem of the workload."""
    import numpy as np
    from oracle import oracle
    oracle.set_threads()                 # all host cores (torchrun exports OMP_NUM_THREADS=1)
    rng = np.random.RandomState(SEED)
    centres = rng.randn(64, d).astype(np.float32)
    x = rng.randn(n_s, d).astype(np.float32) + centres[rng.randint(0, 64, n_s)]
    y = x + sigma * rng.randn(n_s, d).astype(np.float32)
    x = oracle.bf16_round(oracle.normalize_rows(x))
    y = oracle.bf16_round(oracle.normalize_rows(y))
    for _ in range(warmup):
        oracle.align_eval(x, y, True, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = oracle.align_eval(x, y, True, k)
        oracle.metrics(out["rank_l2r"])
        oracle.metrics(out["rank_r2l"])
    dt = (time.perf_counter() - t0) / steps
    return {"value": n_s * n_s / dt, "unit": UNIT, "cores": oracle.max_threads(), "kind": "port",
            "sample": f"{n_s} x {n_s} pair sub-problem of the workload (D={d}, k={k}), {steps} timed passes, "
                      f"{dt * 1e3:.0f} ms each; oracle/snag_oracle.c with OpenMP on all host threads",
            "ms_per_step": dt * 1e3}


def reference_steps_sample(n_s: int, d: int, k: int, sigma: float):
    """The reference's literal op sequence (torch.mm, topk, per-row torch.sort + .item(), main.py:385-429) restated with
    the same torch CPU calls, timed once on a smaller sample — reported beside the port for context."""
    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.RandomState(SEED)
    x = torch.from_numpy(rng.randn(n_s, d).astype(np.float32))
    y = x + sigma * torch.from_numpy(rng.randn(n_s, d).astype(np.float32))
    x, y = torch.nn.functional.normalize(x), torch.nn.functional.normalize(y)
    t0 = time.perf_counter()
    x_norm = (x ** 2).sum(1).view(-1, 1)
    y_norm = (y ** 2).sum(1).view(1, -1)
    distance = torch.clamp(x_norm + y_norm - 2.0 * torch.mm(x, torch.transpose(y, 0, 1)), 0.0, np.inf)
    sim = 1 - distance
    nv1 = torch.mean(torch.topk(sim, k)[0], 1)
    nv2 = torch.mean(torch.topk(sim.t(), k)[0], 1)
    distance = 1 - ((2 * sim.t() - nv1).t() - nv2)
    mrr = 0.0
    for idx in range(n_s):
        _, indices = torch.sort(distance[idx, :], descending=False)
        rank = (indices == idx).nonzero(as_tuple=False).squeeze().item()
        mrr += 1.0 / (rank + 1)
    for idx in range(n_s):
        _, indices = torch.sort(distance[:, idx], descending=False)
        rank = (indices == idx).nonzero(as_tuple=False).squeeze().item()
        mrr += 1.0 / (rank + 1)
    dt = time.perf_counter() - t0
    return {"value": n_s * n_s / dt, "unit": UNIT, "threads": torch.get_num_threads(), "sample_n": n_s, "seconds": dt}


def run_reference(args, name, n, d, k, sigma, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                      # under torchrun only rank 0 times the CPU arm; the others exit 0 without work
    n_s = min(CPU_SAMPLE_N, n)
    cb = cpu_port_sample(n_s, d, k, sigma, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp64-accumulated dots)", "data": "synthetic",
        "config": {"workload": name, "description": desc, "n_pairs": n, "width": d, "csls_k": k,
                   "sampled": f"{n_s} x {n_s} sub-problem per step"},
        "cpu_baseline": {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ================================================================================================ GPU arm
def run_snag(args, name, n, d, k, sigma, desc):
    import torch
    import torch.distributed as dist
    from snag_b200 import evaluate, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    peaks = load_peaks()
    W, K = max(3, args.warmup), max(1, args.steps)

    emb, left, right = synth_tables(n, d, sigma, dev)
    dpad = ops.round_up(d, 64)
    sweep_events = []                       # (name, start, end) per fused-sweep launch inside the timed region

    def timed_sweeps(enable):
        ops.SWEEP_EVENT_SINK = sweep_events if enable else None

    def step_device():
        X, xn = ops.prep_bf16(emb, left, True)
        Y, yn = ops.prep_bf16(emb, right, True)
        return evaluate.align_ranks(X, Y, xn, yn, n, k, True, False, group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing -> value, roofline
    res = None
    for _ in range(W):
        res = step_device()
    barrier()
    timed_sweeps(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(K):
            res = step_device()
        e1.record()
        barrier()
    timed_sweeps(False)
    ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    launches_per_step = res.launches + 2
    # per-kernel durations of the fused sweeps (this rank), algorithmic flops = 2 * rows * cols * D per launch
    kern = {}
    for nm, a, b, rows, cols, _depth in sweep_events:
        kern.setdefault(nm, {"ms": [], "flops": 2.0 * rows * cols * d})["ms"].append(a.elapsed_time(b))
    kstats = {nm: {"launches": len(v["ms"]), "avg_ms": sum(v["ms"]) / len(v["ms"]),
                   "tflops": v["flops"] / (sum(v["ms"]) / len(v["ms"])) / 1e9} for nm, v in kern.items()}
    dom = max(kstats, key=lambda nm: kstats[nm]["avg_ms"] * kstats[nm]["launches"])
    # full sweeps over S executed per step: 3 on the classic path; 2 + m/n with the two-sweep CSLS path
    plan2 = evaluate.two_sweep_plan(n, k)
    sweeps = 3.0 if plan2 is None else 2.0 + plan2[0] / n
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{dom}:{name}:{world}")
    roofline = {"bound": "tensor", "kernel": dom, "achieved": kstats[dom]["tflops"], "peak": peaks["tensor_sustained"],
                "unit": "TFLOP/s", "frac": kstats[dom]["tflops"] / peaks["tensor_sustained"], "traffic": traffic,
                "peak_source": peaks["source"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": kstats[dom]["tflops"] / peaks["tensor_burst"],
                "algorithmic_flops_per_launch": kern[dom]["flops"], "kernels": kstats,
                "sweeps_per_step": sweeps, "executed_tflops_whole_step": sweeps * 2.0 * n * n * d / world / ms_per_step / 1e9}

    if args.profile_run:
        if rank == 0:
            print(json.dumps({"profile_run": True, "workload": name, "ms_per_step": ms_per_step, "kernels": kstats}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------------------------------------------------------- end to end from pinned host memory -> e2e
    c0, c1 = (0, n) if world == 1 else (rank * ((n + world - 1) // world), min(n, (rank + 1) * ((n + world - 1) // world)))
    host = torch.empty((2, max(c1 - c0, 1), d), dtype=torch.float32).pin_memory()
    host[0, :c1 - c0].copy_(emb[c0:c1])
    host[1, :c1 - c0].copy_(emb[n + c0:n + c1])
    del emb
    torch.cuda.empty_cache()

    def step_e2e():
        out = evaluate.evaluate_alignment_host(host[0, :c1 - c0], host[1, :c1 - c0], n, c0, csls=True, csls_k=k, group=group)
        return out

    n_e2e = 2 if n >= 500_000 else 5
    out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        out = step_e2e()
    barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": n * n / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(2 * (c1 - c0) * d * 4) * 1,
           "d2h_bytes_per_step": int(2 * n * 4), "ms_per_step": float(dt.item()) * 1e3, "steps": n_e2e,
           "note": "per rank: H2D of its slice of both fp32 tables from pinned memory (+ NVLink all-gather of the bf16 "
                   "operands when sharded), evaluation, D2H of both rank vectors, host Hits/MR/MRR"}
    metrics = out["l2r"]

    if rank == 0:
        n_s = min(CPU_SAMPLE_N, n)
        cb = cpu_port_sample(n_s, d, k, sigma, 3, 1)
        extra_ref = reference_steps_sample(min(2048, n), d, k, sigma)
        line = {
            "metric": METRIC, "value": n * n / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 operands, fp32 accumulate (tcgen05 kind::f16), fp32 CSLS chain, int32 ranks",
            "data": "synthetic",
            "config": {"workload": name, "description": desc, "n_pairs": n, "width": d, "padded_width": dpad, "csls_k": k,
                       "sigma": sigma, "seed": SEED, "parallelism": f"targets sharded over {world} rank(s)",
                       "l2": "inputs larger than L2 (no flush needed)" if 2 * n * dpad * 2 > 200e6 else
                             "inputs smaller than L2; every step re-reads the fp32 table and rewrites the operands "
                             f"({2 * 2 * n * d * 4 / 1e6:.0f} MB), which exceeds and evicts L2"},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * K,
            "roofline": roofline,
            "cpu_baseline": {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")},
            "reference_literal_steps": extra_ref,
            "quality": {"hits@1_l2r": float(metrics.acc[0]), "hits@10_l2r": float(metrics.acc[1]), "mrr_l2r": metrics.mrr},
            "algorithmic_tflops": 2.0 * n * n * d / (ms_per_step * 1e-3) / 1e12,
            "rank_sweep": res.info.get("rank_sweep", {}),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ================================================================================================ training slice
def _train_tables(B, M, dm, device, seed=SEED):
    """Stand-ins for the encoder outputs of one step (model/SNAG.py:101-102): M modality embeddings and M hidden-state
    embeddings [N, dm], two joint embeddings [N, M*dm], modality weights [N, M]; N = 2B + 1000 entities, B random links."""
    import numpy as np
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    n_ent = 2 * B + 1000
    mk = lambda w: torch.randn((n_ent, w), generator=g, device=device).requires_grad_(True)
    present = [True] * M + [False] * (6 - M)              # (gph, rel, att, img, name, char): surface streams last
    streams = [mk(dm) if p else None for p in present]
    hidden = [mk(dm) if p else None for p in present]
    joint, joint_fz = mk(M * dm), mk(M * dm)
    wn = torch.softmax(torch.randn((n_ent, 6), generator=g, device=device), 1).requires_grad_(True)
    rng = np.random.RandomState(seed)
    links = np.stack([rng.permutation(n_ent // 2)[:B], n_ent // 2 + rng.permutation(n_ent // 2)[:B]], 1).astype(np.int32)
    leaves = [t for t in streams + hidden + [joint, joint_fz, wn] if t is not None]
    return streams, hidden, joint, joint_fz, wn, links, leaves


def cpu_icl_sample(B_s, M, dm, steps, B_full=None):
    """The reference's icl_loss op sequence (model/SNAG_loss.py:58-128: normalise, 4 matmuls, -1e9 self mask, concat,
    log_softmax against one-hot labels, weighted mean) restated with the same torch CPU calls, forward + backward, for
    the 2 + 2M calls of one step at batch B_s, all host threads."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(SEED)
    n_ent = 2 * B_s + 100
    links = torch.stack([torch.randperm(n_ent // 2, generator=g)[:B_s], n_ent // 2 + torch.randperm(n_ent // 2, generator=g)[:B_s]], 1)

    def icl(emb, w):
        z = F.normalize(emb, dim=1)
        a, b = z[links[:, 0]], z[links[:, 1]]
        eye = torch.eye(B_s) * 1e9
        lab = F.one_hot(torch.arange(B_s), 2 * B_s).float()
        la = torch.cat([a @ b.t(), a @ a.t() - eye], 1) / 0.1
        lb = torch.cat([b @ a.t(), b @ b.t() - eye], 1) / 0.1
        wt = torch.ones(B_s) if w is None else torch.min(w[links[:, 0]], w[links[:, 1]])
        xa = -(wt * (lab * F.log_softmax(la, 1)).sum(1)).sum() / B_s
        xb = -(wt * (lab * F.log_softmax(lb, 1)).sum(1)).sum() / B_s
        return 0.5 * xa + 0.5 * xb

    embs = [torch.randn((n_ent, dm), generator=g).requires_grad_(True) for _ in range(2 * M)]
    joints = [torch.randn((n_ent, M * dm), generator=g).requires_grad_(True) for _ in range(2)]
    w = torch.rand((n_ent,), generator=g)
    t0 = time.perf_counter()
    for _ in range(steps):
        tot = sum(icl(e, w if i < M else None) for i, e in enumerate(embs)) + sum(icl(j, None) for j in joints)
        tot.backward()
    dt = (time.perf_counter() - t0) / steps
    scale = 1.0 if not B_full else (B_s / float(B_full)) ** 2          # every term of the step is O(B^2 D)
    return {"value": scale / dt, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one step at batch {B_s} ({2 + 2 * M} icl_loss calls fwd+bwd, the reference's torch op sequence on CPU "
                      f"tensors), {steps} timed step(s), {dt * 1e3:.0f} ms each, scaled by (B_s/B)^2 = {scale:.4g} to the "
                      f"workload's batch (the step is O(B^2 D))"}


def run_train(args, name):
    import numpy as np
    import torch
    import torch.distributed as dist
    from snag_b200 import loss as sloss, ops

    B, M, dm, desc = TRAIN_WORKLOADS[name]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            cb = cpu_icl_sample(min(B, 2048), M, dm, max(1, args.steps), B)
            print(json.dumps({"impl": "reference", "metric": TRAIN_METRIC, "value": cb["value"], "unit": "steps/s",
                              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"],
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": name, "description": desc, "sampled": cb["sample"]}, "cpu_baseline": cb,
                              "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}), flush=True)
        return
    if world != args.gpus:
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    peaks = load_peaks()
    W, K = max(3, args.warmup), max(1, args.steps)
    streams, hidden, joint, joint_fz, wn, links, leaves = _train_tables(B, M, dm, dev)
    layer = sloss.SnagLossLayer(tau=0.1, ab_weight=0.5, awloss=True).to(dev)
    if group is not None:
        layer.distribute(group, grads="gather")
    links_pinned = torch.from_numpy(links).pin_memory()

    links_dev = links_pinned.to(dev)
    use_graph = not args.no_graph and (world == 1 or args.graph_multi)
    eager_ms = None

    def eager_step():
        for t in leaves:
            t.grad = None
        loss = layer(streams, hidden, joint, joint_fz, links_dev, wn)
        loss.backward()
        return loss

    if use_graph:
        # the eager step first, for the record: at the reference's batch sizes it is launch bound
        for _ in range(3):
            eager_step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            eager_step()
        b.record()
        torch.cuda.synchronize()
        eager_ms = a.elapsed_time(b) / K
        from snag_b200.graphs import GraphedStep
        graphed = GraphedStep(lambda: layer(streams, hidden, joint, joint_fz, links_dev, wn), leaves)

    def step(from_host):
        if from_host:                                       # this step's batch arrives from pinned host memory
            links_dev.copy_(links_pinned, non_blocking=True)
        loss = graphed() if use_graph else eager_step()
        return loss.item() if from_host else loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(False)
    barrier()
    events = []
    if not use_graph:
        ops.SWEEP_EVENT_SINK = events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(K):
            step(False)
        e1.record()
        barrier()
    ops.SWEEP_EVENT_SINK = events
    ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    if use_graph:                     # per-kernel CUDA events cannot be recorded into a replayed graph: time the sweeps of
        for _ in range(K):            # K eager steps of the same work instead (same kernels, same arguments)
            eager_step()
        torch.cuda.synchronize()
    ops.SWEEP_EVENT_SINK = None
    kern = {}
    for nm, a, b, rows, cols, depth in events:
        e = kern.setdefault(nm, {"ms": 0.0, "flops": 0.0, "launches": 0})
        e["ms"] += a.elapsed_time(b)
        e["flops"] += 2.0 * rows * cols * depth
        e["launches"] += 1
    kstats = {nm: {"launches_per_step": v["launches"] // K, "ms_per_step": v["ms"] / K, "tflops": v["flops"] / v["ms"] / 1e9}
              for nm, v in kern.items()}
    dom = max(kstats, key=lambda nm: kstats[nm]["ms_per_step"])
    n_calls = 2 + 2 * M
    d_sum = 2 * M * dm + 2 * M * dm                        # sum of the contraction widths over the calls
    alg_fwd = 6.0 * B * B * d_sum                          # 3 distinct B x B x D contractions per call
    alg_bwd = 14.0 * B * B * d_sum
    roofline = {"bound": "tensor", "kernel": dom, "achieved": kstats[dom]["tflops"], "peak": peaks["tensor_burst"],
                "unit": "TFLOP/s", "frac": kstats[dom]["tflops"] / peaks["tensor_burst"], "traffic": None,
                "peak_source": peaks["source"] + ", bf16_tflops (burst: launches of a few ms, timed alone with CUDA events)",
                "kernels": kstats, "note": "flops counted on the padded contraction width the kernel executes; "
                "sim_kernel<EpiWrite> is the gradient GEMM dX = G.[other;this] (contraction width 2*Bp)",
                "algorithmic_flops_per_step": alg_fwd + alg_bwd,
                "algorithmic_tflops_whole_step": (alg_fwd + alg_bwd) / world / ms_per_step / 1e9}
    # end to end: the step's input (the batch of links) comes from pinned host memory, the loss goes back to the host
    step(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        lv = step(True)
    barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        cb = cpu_icl_sample(min(B, 2048), M, dm, 1, B)
        line = {"metric": TRAIN_METRIC, "value": 1e3 / ms_per_step, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16 operands, fp32 accumulate (tcgen05 kind::f16), fp32 softmax statistics, bf16 dL/dlogits",
                "data": "synthetic",
                "config": {"workload": name, "description": desc, "batch": B, "modalities": M, "width": dm, "tau": 0.1,
                           "icl_calls_per_step": n_calls, "parallelism": f"anchors sharded over {world} rank(s)",
                           "cuda_graph": use_graph, "eager_ms_per_step": eager_ms,
                           "l2": "every step rewrites the bf16 operands and dL/dlogits (> L2) between launches"},
                "clocks": clocks.summary(),
                "e2e": {"value": 1.0 / float(dt.item()), "unit": "steps/s", "h2d_bytes_per_step": int(links_pinned.numel() * 4),
                        "d2h_bytes_per_step": 4, "ms_per_step": float(dt.item()) * 1e3, "loss": lv},
                "gpu_launches": sum(v["launches"] for v in kern.values()) + 4 * n_calls * K,
                "roofline": roofline, "cpu_baseline": cb}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("SNAG_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS) + sorted(TRAIN_WORKLOADS))
    ap.add_argument("--impl", default="snag", choices=["snag", "reference"])
    ap.add_argument("--graph-multi", action="store_true", help="training slice: capture the NCCL exchanges in the graph too (N > 1)")
    ap.add_argument("--no-graph", action="store_true", help="training slice: run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-run", action="store_true",
                    help="only the device-resident timed region (for runs under ncu); e2e and CPU legs are skipped")
    args = ap.parse_args()
    if args.workload in TRAIN_WORKLOADS:
        run_train(args, args.workload)
        return
    n, d, k, sigma, desc = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, n, d, k, sigma, desc)
    else:
        run_snag(args, args.workload, n, d, k, sigma, desc)


if __name__ == "__main__":
    main()
