"""Kernel-level breakdown of one 1M x 1M evaluation (torch.profiler): the sweeps and everything around them."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from snag_b200 import evaluate, ops


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d, k, sigma = 1200, 10, (6.0 if n >= 200_000 else 8.0)
    dev = torch.device("cuda:0")
    emb, left, right = bench.synth_tables(n, d, sigma, dev)

    def step():
        X, xn = ops.prep_bf16(emb, left, True)
        Y, yn = ops.prep_bf16(emb, right, True)
        return evaluate.align_ranks(X, Y, xn, yn, n, k, True, False, None)

    reps = 1 if n >= 200_000 else 10
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t / reps, e.count // reps, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"# one evaluation of {n} x {n}, D={d}: kernels by CUDA time (us, launches); total {tot / 1e3:.2f} ms")
    for t, c, kk in rows[:45]:
        print(f"{t:12.1f} us  x{c:<4d} {100 * t / tot:5.1f}%  {kk[:120]}")


if __name__ == "__main__":
    main()
