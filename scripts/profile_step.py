"""Kernel-level breakdown of one eager loss-layer step (torch.profiler / kineto; nsys is not in the image): which
kernels — ours and torch's glue — make up the step. Writes a table sorted by total CUDA time.

    python scripts/profile_step.py c5_train > gpurun_out/step_profile_c5.txt
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from snag_b200 import loss as sloss


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c5_train"
    B, M, dm, _ = bench.TRAIN_WORKLOADS[name]
    dev = torch.device("cuda:0")
    streams, hidden, joint, joint_fz, wn, links, leaves = bench._train_tables(B, M, dm, dev)
    layer = sloss.SnagLossLayer(tau=0.1, ab_weight=0.5).to(dev)
    links_dev = torch.from_numpy(links).to(dev)

    def step():
        for t in leaves:
            t.grad = None
        loss = layer(streams, hidden, joint, joint_fz, links_dev, wn)
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t / 3.0, e.count // 3, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"# {name}: one eager step, kernels by CUDA time (us per step, launches per step); total {tot / 1e3:.2f} ms")
    for t, c, k in rows[:40]:
        print(f"{t:10.1f} us  x{c:<4d} {100 * t / tot:5.1f}%  {k[:110]}")


if __name__ == "__main__":
    main()
