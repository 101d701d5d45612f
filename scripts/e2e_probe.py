"""Dev probe: timeline of the streamed host entry point at n x n (default 1M), D = 1200."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from snag_b200 import evaluate, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d, k, sigma = 1200, 10, 6.0
dev = torch.device("cuda", 0)
emb, left, right = bench.synth_tables(n, d, sigma, dev)
host = torch.empty((2, n, d), dtype=torch.float32).pin_memory()
host[0].copy_(emb[:n])
host[1].copy_(emb[n:2 * n])
del emb
torch.cuda.empty_cache()
print("pinned:", host[0].is_pinned(), host[0, :n].is_pinned(), flush=True)
m = evaluate.two_sweep_plan(n, k)[0]

# raw H2D bandwidth
buf = torch.empty((n, d), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
buf.copy_(host[0], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"H2D one table: {dt * 1e3:.1f} ms = {n * d * 4 / dt / 1e9:.1f} GB/s", flush=True)
del buf

# host gather
t0 = time.perf_counter()
stage = evaluate._pinned((m, d), ("ysamp", m, d))
t1 = time.perf_counter()
torch.index_select(host[1], 0, evaluate._sample_rows_host(n, m, dev)[1], out=stage)
t2 = time.perf_counter()
print(f"pinned staging alloc {1e3 * (t1 - t0):.1f} ms, index_select of {m} rows {1e3 * (t2 - t1):.1f} ms", flush=True)

for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tl = {}
    X, Y, xn, yn, pre, launches = evaluate._stream_in_with_prepasses(host[0], host[1], n, k, True, dev, m, timeline=tl)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("   timeline (ms after start):", {kk: round(tl["start"].elapsed_time(v), 1) for kk, v in tl.items() if kk != "start"}, flush=True)
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, True, False, None, pre=pre)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"[{it}] stream-in enqueue {1e3 * (t1 - t0):.1f} ms, until landed+prepassed {1e3 * (t2 - t0):.1f} ms, "
          f"align_ranks(pre) {1e3 * (t3 - t2):.1f} ms", flush=True)
    del X, Y, xn, yn, pre, res
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = evaluate.evaluate_alignment_host(host[0], host[1], n, csls=True, csls_k=k, stream_in=False)
    t1 = time.perf_counter()
    o2 = evaluate.evaluate_alignment_host(host[0], host[1], n, csls=True, csls_k=k)
    t2 = time.perf_counter()
    print(f"[{it}] plain {1e3 * (t1 - t0):.1f} ms, streamed {1e3 * (t2 - t1):.1f} ms", flush=True)
