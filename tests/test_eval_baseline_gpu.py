"""GPU parity AT THE BASELINE SIZES: the fused evaluation (through the C ABI) against the oracle on every pair —
ranks of both directions, CSLS neighbourhood means, ground-truth distances and Hits@k / MR / MRR, all bit-exact —
for the evaluation shapes BASELINE.json's configs name (C1 10 500 x 1200 with the default k = 10 and the
scripted k = 3 of run_snag.sh:17, C2 10 500 x 1800, C3 10 277 x 1200) and for one size above
evaluate.TWO_SWEEP_MIN_N at D = 1200, so that the two-sweep path the 1M headline runs (sample pre-passes, per-CTA
candidate streams, column bucketing) meets the oracle and not only the repo's own three-sweep path.
Reference: main.py:385-429, src/utils.py:202-218, 417-435."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import oracle
from snag_b200 import evaluate, ops

pytestmark = pytest.mark.gpu


def _clustered(n, d, sigma, seed):
    """SURVEY 8(d) generator (64 cluster centres, noisy copies as targets), normalised and rounded to bf16 on the host
    so that the oracle and the device see identical values."""
    rng = np.random.RandomState(seed)
    centres = rng.randn(64, d).astype(np.float32)
    x = rng.randn(n, d).astype(np.float32) + centres[rng.randint(0, 64, n)]
    y = x + np.float32(sigma) * rng.randn(n, d).astype(np.float32)
    return oracle.bf16_round(oracle.normalize_rows(x)), oracle.bf16_round(oracle.normalize_rows(y))


def _prep(x, y, dev):
    X, xn = ops.prep_bf16(torch.from_numpy(np.ascontiguousarray(x)).to(dev), None, normalize=False)
    Y, yn = ops.prep_bf16(torch.from_numpy(np.ascontiguousarray(y)).to(dev), None, normalize=False)
    return X, Y, xn, yn


def _assert_equal_to_oracle(res, ref, csls):
    if csls:
        np.testing.assert_array_equal(res.nv1.cpu().numpy(), ref["nv1"])
        np.testing.assert_array_equal(res.nv2.cpu().numpy(), ref["nv2"])
        info = res.info["neighbourhoods"]
        assert info["rows"]["unverified"] == 0 and info["cols"]["unverified"] == 0
    np.testing.assert_array_equal(res.g.cpu().numpy(), ref["g"])
    np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), ref["rank_l2r"])
    np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), ref["rank_r2l"])
    for side, ranks in (("l2r", res.rank_l2r), ("r2l", res.rank_r2l)):
        m = evaluate.metrics_from_ranks(ranks)
        o = oracle.metrics(ref[f"rank_{side}"])
        assert m.mr == o["mr"] and m.mrr == o["mrr"] and np.array_equal(m.acc, o["acc"])


@pytest.mark.parametrize("name,n,d,k,csls", [
    ("c1_k10", 10500, 1200, 10, True),       # configs[0]/[1] DBP15K ja_en-shaped, config.py:65 default k
    ("c1_k3", 10500, 1200, 3, True),         # the scripted --csls_k 3 (run_snag.sh:17)
    ("c2_k10", 10500, 1800, 10, True),       # fr_en + surface: joint width 1800
    ("c3_k10", 10277, 1200, 10, True),       # FBDB15K-shaped
    ("c1_nocsls", 10500, 1200, 10, False),   # --csls off: ranks on the squared distance itself (main.py:392)
])
def test_baseline_config_against_oracle(cuda_device, name, n, d, k, csls):
    oracle.set_threads()
    x, y = _clustered(n, d, 8.0, 3408)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, csls, want_top3=True)
    ref = oracle.align_eval(x, y, csls, k)
    _assert_equal_to_oracle(res, ref, csls)
    hits1 = float(evaluate.metrics_from_ranks(res.rank_l2r).acc[0])
    assert 0.3 < hits1 < 0.999, hits1         # the workload is neither trivial nor noise
    # prediction-file ids (main.py:411): exact wherever the 4th and 5th nearest can be told apart by the tensor cores
    top3 = res.top3_idx.cpu().numpy()
    assert (top3 == ref["top3"]).all(axis=1).mean() > 0.999
    # sharded over 2 and 8 ranks (lockstep-simulated on this GPU): bit-identical to the oracle as well
    for world in (2, 8):
        many = evaluate.simulate_sharded(
            lambda r: evaluate._align_ranks_steps(ops, X, Y, xn, yn, n, k, csls, False, world, r), world)
        _assert_equal_to_oracle(many[world - 1], ref, csls)


def test_two_sweep_size_against_oracle(cuda_device):
    """n just above TWO_SWEEP_MIN_N at the headline width D = 1200, k = 10: the paths bench.py's c4 workloads run — the
    one-pass evaluation (default) and the two-sweep evaluation — against the oracle on every pair."""
    import psutil
    oracle.set_threads()
    n, d, k = evaluate.TWO_SWEEP_MIN_N + 1, 1200, 10
    assert evaluate.two_sweep_plan(n, k) is not None
    x, y = _clustered(n, d, 8.0, 3409)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    # the oracle: materialised when the n x n fp32 matrix fits comfortably in host memory (one pass over the dot
    # products), streaming otherwise (two passes)
    if psutil.virtual_memory().available > 3 * 4 * n * n:
        ref = oracle.align_eval(x, y, True, k)
    else:
        ref = oracle.align_eval_stream(x, y, True, k, 2048)
    for one_pass, kernel in ((True, "sim_kernel<EpiOnePass>"), (False, "sim_kernel<EpiRowColTopK>")):
        sweeps = []
        ops.SWEEP_EVENT_SINK = sweeps
        try:
            res = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=one_pass)
        finally:
            ops.SWEEP_EVENT_SINK = None
        kinds = [nm for nm, *_ in sweeps]
        assert kernel in kinds                           # one sweep for both CSLS directions
        _assert_equal_to_oracle(res, ref, True)
        if one_pass:
            info = res.info["one_pass"]
            assert info["fallback"] is None and "sim_kernel<EpiRank>" not in kinds, info    # no second sweep
            assert 0 < info["streamed"] < 2e-3 * n * n, info
            print(f"\n[one-pass] n={n}: streamed {info['streamed']} ({info['streamed'] / n / n:.2e} of the pairs), "
                  f"deferred {info['deferred']}, failed guesses rows/cols {info['failed_rows']}/{info['failed_cols']}")
        # the same through 4 simulated ranks
        many = evaluate.simulate_sharded(
            lambda r: evaluate._align_ranks_steps(ops, X, Y, xn, yn, n, k, True, False, 4, r, one_pass=one_pass), 4)
        _assert_equal_to_oracle(many[1], ref, True)
        if one_pass:
            assert all(mr.info["one_pass"]["fallback"] is None for mr in many)


def test_streamed_host_entry_point_equals_resident(cuda_device):
    """evaluate_alignment_host with the chunked, overlapped transfer (prologue and sample pre-passes run on the chunks as
    they arrive) returns the same ranks and neighbourhood means, bit for bit, as the same tables evaluated after one
    plain copy — with a chunk size that does not divide n."""
    n, d, k = evaluate.TWO_SWEEP_MIN_N + 4321, 320, 10
    rng = np.random.RandomState(11)
    centres = rng.randn(64, d).astype(np.float32)
    x = rng.randn(n, d).astype(np.float32) + centres[rng.randint(0, 64, n)]
    y = x + np.float32(3.0) * rng.randn(n, d).astype(np.float32)
    hx, hy = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    old = evaluate.STREAM_CHUNK_ROWS
    evaluate.STREAM_CHUNK_ROWS = 20000
    try:
        a = evaluate.evaluate_alignment_host(hx, hy, n, csls=True, csls_k=k, device=cuda_device)
    finally:
        evaluate.STREAM_CHUNK_ROWS = old
    b = evaluate.evaluate_alignment_host(hx, hy, n, csls=True, csls_k=k, device=cuda_device, stream_in=False)
    assert a.get("streamed") and not b.get("streamed")
    for key in ("rank_l2r", "rank_r2l", "nv1", "nv2", "g"):
        assert torch.equal(getattr(a["ranks"], key), getattr(b["ranks"], key)), key
    assert a["l2r"].mrr == b["l2r"].mrr and a["r2l"].mrr == b["r2l"].mrr
    assert a["ranks"].info["one_pass"]["fallback"] is None


def test_one_pass_fallbacks_are_exact(cuda_device, monkeypatch):
    """The one-pass evaluation speculates on upper bounds of the neighbourhood means; whatever happens to the guesses the
    ranks must not change: (a) guesses that are far too low -> the failed entities are recounted exhaustively,
    (b) more failures than the exhaustive budget -> classic sweep 2, (c) a rank stream that overflows -> classic sweep 2
    and a larger stream next time. Compared bit for bit with the two-sweep evaluation."""
    n, d, k = evaluate.TWO_SWEEP_MIN_N + 777, 320, 10
    g = torch.Generator(device="cuda").manual_seed(12)
    centres = torch.randn((64, d), generator=g, device="cuda")
    x = torch.randn((n, d), generator=g, device="cuda") + centres[torch.randint(0, 64, (n,), generator=g, device="cuda")]
    y = x + 3.0 * torch.randn((n, d), generator=g, device="cuda")
    X, xn = ops.prep_bf16(x, None, True)
    Y, yn = ops.prep_bf16(y, None, True)
    del x, y
    ref = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=False)
    assert ref.info["one_pass"] is None

    def same(res):
        return (torch.equal(res.rank_l2r, ref.rank_l2r) and torch.equal(res.rank_r2l, ref.rank_r2l) and
                torch.equal(res.nv1, ref.nv1) and torch.equal(res.nv2, ref.nv2) and torch.equal(res.g, ref.g))

    a = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=True)
    assert same(a) and a.info["one_pass"]["fallback"] is None, a.info["one_pass"]
    # (a) no extrapolation at all: the guess is the sample's own mean -> many entities fail and are recounted by
    # tensor-core sweeps over the gathered sub-panels
    monkeypatch.setattr(evaluate, "ONE_PASS_GAMMA", -0.02)
    monkeypatch.setattr(evaluate, "ONE_PASS_RECOUNT_MAX_FRAC", 10.0)
    b = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=True)
    ib = b.info["one_pass"]
    assert ib["fallback"] is None and ib["failed_rows"] > 32 and ib["failed_cols"] > 32, ib
    assert same(b)
    # ... or, when they are few (here: forced), exhaustively with fp64 dot products
    monkeypatch.setattr(evaluate, "ONE_PASS_EXHAUSTIVE_MAX", 1 << 30)
    b2 = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=True)
    assert b2.info["one_pass"]["fallback"] is None and same(b2)
    monkeypatch.setattr(evaluate, "ONE_PASS_EXHAUSTIVE_MAX", 32)
    # sharded (3 simulated ranks) with failing guesses
    many = evaluate.simulate_sharded(
        lambda r: evaluate._align_ranks_steps(ops, X, Y, xn, yn, n, k, True, False, 3, r, one_pass=True), 3)
    assert all(same(mr) for mr in many)
    # (b) recounting would cost more than sweep 2 -> sweep 2
    monkeypatch.setattr(evaluate, "ONE_PASS_RECOUNT_MAX_FRAC", 0.0)
    c = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=True)
    assert c.info["one_pass"]["fallback"] == "too many failed guesses" and same(c)
    # (c) stream overflow -> sweep 2, and the capacity is raised for the next evaluations of this shape
    monkeypatch.setattr(evaluate, "ONE_PASS_GAMMA", 1.5)
    monkeypatch.setattr(evaluate, "ONE_PASS_RECOUNT_MAX_FRAC", 0.8)
    evaluate._ONE_PASS_CAP.clear()
    evaluate._ONE_PASS_CAP[(n, n, k, X.shape[1])] = 64               # far too small on purpose
    e = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=True)
    assert e.info["one_pass"]["fallback"] == "rank stream overflow" and same(e), e.info["one_pass"]
    for _ in range(8):
        f = evaluate.align_ranks(X, Y, xn, yn, n, k, True, one_pass=True)
        assert same(f)
        if f.info["one_pass"]["fallback"] is None:
            break
    assert f.info["one_pass"]["fallback"] is None, f.info["one_pass"]
    evaluate._ONE_PASS_CAP.clear()


def _adversarial_rows(n, d, seed):
    """Unit rows built to stress the tensor core's fp32 accumulation against the canonical fp64 one: dense clustered
    rows (typical), all-positive rows (no cancellation: the largest partial sums), constant rows, rows whose mass sits
    in a few coordinates, sign-alternating copies (s = sum |x_k y_k| cancels to ~0) — and the SAME rows on both sides,
    so that every kind also meets its exact duplicate (s = ||x||^2 ~ 1) and its near-duplicates."""
    rng = np.random.RandomState(seed)
    q = n // 8
    centres = rng.randn(16, d).astype(np.float32)
    dense = rng.randn(2 * q, d).astype(np.float32) + centres[rng.randint(0, 16, 2 * q)]
    pos = np.abs(rng.randn(q, d)).astype(np.float32)
    const = np.ones((q, d), np.float32) * (1 + 1e-3 * rng.rand(q, 1).astype(np.float32))
    spiky = (rng.randn(q, d) * (rng.rand(q, d) < 0.02)).astype(np.float32) + 1e-3 * rng.randn(q, d).astype(np.float32)
    alt = pos * np.where(np.arange(d) % 2 == 0, 1.0, -1.0).astype(np.float32)
    near = dense[:q] + 1e-3 * rng.randn(q, d).astype(np.float32)              # near-duplicates of the first dense rows
    rest = rng.randn(n - 7 * q, d).astype(np.float32)
    x = np.concatenate([dense, pos, const, spiky, alt, near, rest], 0)
    return oracle.bf16_round(oracle.normalize_rows(x))


@pytest.mark.parametrize("d", [1200, 1856])
def test_band_epsilon_covers_tensor_core_error_on_1e9_pairs(cuda_device, d):
    """ops.tc_margin (the rank sweep's deferral band and the neighbourhood verification) rests on a bound for
    |s_tensor_core - s_canonical| on unit rows (ops.TC_DOT_ERR_PER_K x Dpad). Measure it
    on 32 768^2 = 1.07e9 pairs per width (D = 1200 -> Dpad 1216, and Dpad 1856, the widest the configs use) that
    include exact duplicates, near-duplicates, all-positive / constant rows (largest partial sums) and sign-alternating
    rows (full cancellation). Comparator: fp64 GEMM rounded once to fp32 — within half an fp32 ulp (3e-8 for |s| <= 1)
    of the oracle's index-order fp64 dot. The band must keep a 4x margin over the worst error seen."""
    n = 32768
    x = _adversarial_rows(n, d, 77)
    Xf = torch.from_numpy(x).to(cuda_device)
    X, xn = ops.prep_bf16(Xf, None, normalize=False)
    assert torch.equal(X[:, :d].float(), Xf)                       # the operands are exactly the host rows
    S = ops.sim_write(X, X, None, None, n, n, 0)                    # tensor-core dot products, fp32 [n, n]
    Xd = Xf.double()
    worst, worst_diag, pairs = 0.0, 0.0, 0
    for r0 in range(0, n, 2048):
        ref = (Xd[r0:r0 + 2048] @ Xd.t()).float()
        err = (S[r0:r0 + 2048] - ref).abs()
        worst = max(worst, float(err.max()))
        worst_diag = max(worst_diag, float(err[torch.arange(err.shape[0]), r0 + torch.arange(err.shape[0])].max()))
        pairs += err.numel()
    assert pairs >= 1_000_000_000
    # spot-check the comparator itself against the oracle's canonical dot on a few thousand pairs
    rng = np.random.RandomState(1)
    ri, ci = rng.randint(0, n, 4096).astype(np.int32), rng.randint(0, n, 4096).astype(np.int32)
    ri[:1024] = ci[:1024]                                          # include diagonal (duplicate) pairs
    canon = ops.pairs_dot(X, X, torch.from_numpy(ri).to(cuda_device), torch.from_numpy(ci).to(cuda_device))
    ref = (Xd[torch.from_numpy(ri).long().to(cuda_device)] * Xd[torch.from_numpy(ci).long().to(cuda_device)]).sum(1).float()
    assert float((canon - ref).abs().max()) <= 6e-8
    print(f"\n[band] D={d}: max |s_tc - s_fp64| over {pairs:.3e} pairs = {worst:.3e} (diagonal: {worst_diag:.3e}); "
          f"model {ops.TC_DOT_ERR_PER_K * X.shape[1]:.3e}, margin {ops.tc_margin(X.shape[1]):.3e}")
    assert worst <= ops.TC_DOT_ERR_PER_K * X.shape[1], worst             # the model bounds what was measured
    assert 4 * worst <= ops.tc_margin(X.shape[1]), worst                 # and the margins keep 4x over it
