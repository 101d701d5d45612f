"""GPU parity: the fused alignment evaluation (through the C ABI) against the reference's golden vectors, the
oracle on seeded inputs, and size-independent properties at the BASELINE sizes. Ranks / hits / top-3 ids are
compared bit-exactly; CSLS neighbourhood means and distances within 1e-6 (tensor-core vs fp64 accumulation
order of the dot products, everything else in the chain is the same single fp32 op)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import oracle
from snag_b200 import _lib, evaluate, ops
from tests.conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _prep(x, y, dev):
    X, xn = ops.prep_bf16(torch.from_numpy(np.ascontiguousarray(x)).to(dev), None, normalize=False)
    Y, yn = ops.prep_bf16(torch.from_numpy(np.ascontiguousarray(y)).to(dev), None, normalize=False)
    return X, Y, xn, yn


def _clustered(n, d, sigma, seed):
    rng = np.random.RandomState(seed)
    centres = rng.randn(64, d).astype(np.float32)
    x = rng.randn(n, d).astype(np.float32) + centres[rng.randint(0, 64, n)]
    y = x + sigma * rng.randn(n, d).astype(np.float32)
    return oracle.bf16_round(oracle.normalize_rows(x)), oracle.bf16_round(oracle.normalize_rows(y))


@pytest.mark.parametrize("name", golden_names("eval_"))
def test_golden_vectors(cuda_device, name):
    fx = load_golden(name)
    k, csls = int(fx["k"]), bool(fx["csls"])
    X, Y, xn, yn = _prep(fx["x"], fx["y"], cuda_device)
    n = fx["x"].shape[0]
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, csls, want_top3=True)
    np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), fx["rank_l2r"])
    np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), fx["rank_r2l"])
    np.testing.assert_array_equal(res.top3_idx.cpu().numpy(), fx["top3"])
    np.testing.assert_allclose(res.g.cpu().numpy(), fx["g"], rtol=0, atol=2e-6)
    if csls:
        np.testing.assert_allclose(res.nv1.cpu().numpy(), fx["nv1"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(res.nv2.cpu().numpy(), fx["nv2"], rtol=0, atol=1e-6)
        # (the goldens come from the reference's own torch.mm, whose accumulation order is the library's: 1e-6;
        #  against the oracle's canonical accumulation the same quantities are bit-exact, see test_against_oracle)
    for side, ranks in (("l2r", res.rank_l2r), ("r2l", res.rank_r2l)):
        m = evaluate.metrics_from_ranks(ranks)
        np.testing.assert_array_equal(m.acc, fx[f"acc_{side}"])
        assert m.mr == float(fx[f"mr_{side}"]) and m.mrr == float(fx[f"mrr_{side}"])


def test_dyadic_fixture_is_bit_exact_everywhere(cuda_device):
    """Dyadic-rational inputs make every product and partial sum exact in fp32, so the tensor-core path must agree
    with the reference on every float, not only on the ranks — and the fixture forces exact ties."""
    fx = load_golden("eval_ties_dyadic_k4")
    X, Y, xn, yn = _prep(fx["x"], fx["y"], cuda_device)
    res = evaluate.align_ranks(X, Y, xn, yn, fx["x"].shape[0], 4, True)
    np.testing.assert_array_equal(res.nv1.cpu().numpy(), fx["nv1"])
    np.testing.assert_array_equal(res.nv2.cpu().numpy(), fx["nv2"])
    np.testing.assert_array_equal(res.g.cpu().numpy(), fx["g"])


@pytest.mark.parametrize("n,d,k,csls,sigma", [(2048, 1200, 10, True, 8.0), (3000, 1800, 3, True, 8.0),
                                              (1000, 300, 10, False, 3.0), (1531, 96, 16, True, 2.0)])
def test_against_oracle(cuda_device, n, d, k, csls, sigma):
    x, y = _clustered(n, d, sigma, 3408)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, csls, want_top3=True)
    ref = oracle.align_eval(x, y, csls, k, want_dist=True)
    np.testing.assert_array_equal(xn.cpu().numpy(), oracle.norm2(x))
    # Every quantity that leaves the evaluation is computed from canonically accumulated dot products (neighbourhood
    # candidates and near-ties are re-scored in fp64 index order): bit-exact on EVERY pair, ambiguous or not.
    if csls:
        np.testing.assert_array_equal(res.nv1.cpu().numpy(), ref["nv1"])
        np.testing.assert_array_equal(res.nv2.cpu().numpy(), ref["nv2"])
        assert res.info["neighbourhoods"]["rows"]["unverified"] == 0 and res.info["neighbourhoods"]["cols"]["unverified"] == 0
    np.testing.assert_array_equal(res.g.cpu().numpy(), ref["g"])
    np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), ref["rank_l2r"])
    np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), ref["rank_r2l"])
    # top-3 ids: the four nearest by tensor-core score are re-scored canonically; exact unless the 4th and 5th
    # nearest are closer than the tensor core can tell apart
    dist = ref["dist"]
    srt = np.sort(dist, 1)
    clear = srt[:, 4] - srt[:, 3] > 2e-5
    assert clear.mean() > 0.99
    np.testing.assert_array_equal(res.top3_idx.cpu().numpy()[clear], ref["top3"][clear])
    for side, ranks in (("l2r", res.rank_l2r), ("r2l", res.rank_r2l)):
        m = evaluate.metrics_from_ranks(ranks)
        mo = evaluate.metrics_from_ranks(ref[f"rank_{side}"])
        assert m.mr == mo.mr and m.mrr == mo.mrr and np.array_equal(m.acc, mo.acc)


def test_wide_dynamic_range_rows_stay_bit_exact(cuda_device):
    """The canonical re-scores add the fp64 products across 8 lanes only when no addition can round in ANY order
    (canonical_dot_coop: bound < 2^(39 + Ea + Eb)); rows holding very small elements fail that test and must take the
    index-order loop. Elements scaled down to 2^-12 .. 2^-40 of their size (and exact zeros, and a few bf16 denormals)
    exercise both branches; everything stays bit-identical to the oracle, whose fp64 sums DO round here."""
    rng = np.random.RandomState(77)
    n, d, k = 700, 200, 10
    x, y = _clustered(n, d, 2.0, 5)
    scale = np.float32(2.0) ** -rng.randint(12, 41, size=(n, d)).astype(np.float32)
    pick = rng.rand(n, d) < 0.05
    x = np.where(pick, x * scale, x).astype(np.float32)
    y = np.where(rng.rand(n, d) < 0.05, y * scale[::-1], y).astype(np.float32)
    x[rng.rand(n, d) < 0.02] = 0.0
    y[3] = 0.0                                                   # a row of zeros
    x[5, :4] = np.float32(1e-39)                                 # below bf16's normal range
    x, y = oracle.bf16_round(x), oracle.bf16_round(y)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, True)
    ref = oracle.align_eval(x, y, True, k)
    np.testing.assert_array_equal(res.nv1.cpu().numpy(), ref["nv1"])
    np.testing.assert_array_equal(res.nv2.cpu().numpy(), ref["nv2"])
    np.testing.assert_array_equal(res.g.cpu().numpy(), ref["g"])
    np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), ref["rank_l2r"])
    np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), ref["rank_r2l"])


@pytest.mark.parametrize("n,d", [(1500, 1200), (1100, 1800), (2100, 300)])
def test_tensor_core_dot_error(cuda_device, n, d):
    """The deferral band of the rank sweep (ops.tc_margin) must cover the distance between the tensor core's dot
    product and the canonical one (fp64, index order, rounded once) for unit rows: pin it with a 4x safety factor on a few million samples."""
    x, y = _clustered(n, d, 4.0, 11)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    S = ops.sim_write(X, Y, None, None, n, n, 0).cpu().numpy()
    ref = oracle.dot_matrix(x, y)
    assert np.abs(S - ref).max() < ops.tc_margin(X.shape[1]) / 4


@pytest.mark.parametrize("n,d,k,csls,sigma", [(2048, 1200, 10, True, 8.0), (1000, 300, 10, False, 3.0),
                                              (1531, 96, 16, True, 2.0)])
def test_rank_sweep_is_canonical(cuda_device, n, d, k, csls, sigma):
    """Given the neighbourhood means, the rank sweep + band re-score must reproduce the oracle's ranks on EVERY pair,
    near-ties included: elements inside the band are judged with the canonical arithmetic, the others cannot flip."""
    x, y = _clustered(n, d, sigma, 3408)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    ref = oracle.align_eval(x, y, csls, k)
    nv1 = torch.from_numpy(ref["nv1"]).to(cuda_device) if csls else None
    nv2 = torch.from_numpy(ref["nv2"]).to(cuda_device) if csls else None
    g = ops.pair_score(X, Y, n, xn, yn, nv1, nv2, csls)
    np.testing.assert_array_equal(g.cpu().numpy(), ref["g"])
    for exact_chain in (False, True):
        cnt_row = torch.zeros((n,), dtype=torch.int32, device=cuda_device)
        cnt_col = torch.zeros((n,), dtype=torch.int32, device=cuda_device)
        ops.eval_rank(X, Y, xn, yn, nv1, nv2, g, g, 0, 0, n, n, csls, cnt_row, cnt_col, exact_chain=exact_chain)
        if not exact_chain:
            assert ops.LAST_RANK_INFO["mode"] == "band"
            np.testing.assert_array_equal(cnt_row.cpu().numpy(), ref["rank_l2r"])
            np.testing.assert_array_equal(cnt_col.cpu().numpy(), ref["rank_r2l"])
        else:                                  # the in-kernel chain may move a near-tie by one place
            assert np.abs(cnt_row.cpu().numpy() - ref["rank_l2r"]).max() <= 2


def test_rank_band_overflow_retries(cuda_device, monkeypatch):
    """All-equal embeddings put every element inside the band (every distance equals every ground-truth distance): a
    deferral list that is too small is detected, the partial counts are undone and the sweep is re-run with a larger
    list; the result is the stable-sort order (lower ids first)."""
    n, d = 700, 64
    x, _ = _clustered(n, d, 1.0, 5)
    xc = np.tile(x[:1], (n, 1))
    Xc, Yc, xnc, ync = _prep(xc, xc, cuda_device)
    monkeypatch.setattr(ops, "RANK_BAND_MIN_CAP", 16)
    monkeypatch.setattr(ops, "RANK_BAND_PER_ROW", 0)
    for csls in (False, True):
        res = evaluate.align_ranks(Xc, Yc, xnc, ync, n, 3, csls)
        np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), np.arange(n))
        np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), np.arange(n))
        assert ops.LAST_RANK_INFO["deferred"] == n * (n - 1) and ops.LAST_RANK_INFO["cap"] >= n * (n - 1)


def test_rank_chain_fallback_with_top3(cuda_device, monkeypatch):
    """When the deferral list would exceed its maximum size (everything tied: constant embeddings) the in-kernel fp32
    chain with its exact-tie path takes over. That path judges ties on the tensor core's own dot products, so the
    check uses dyadic rows (every product and partial sum exact): ranks and top-3 ids are those of a stable sort."""
    n, d = 300, 64
    xc = np.full((n, d), 0.125, np.float32)                      # ||row||^2 = 1, s = 1 exactly
    Xc, Yc, xnc, ync = _prep(xc, xc, cuda_device)
    monkeypatch.setattr(ops, "RANK_BAND_MIN_CAP", 16)
    monkeypatch.setattr(ops, "RANK_BAND_PER_ROW", 0)
    monkeypatch.setattr(ops, "RANK_BAND_MAX_CAP", 64)
    res = evaluate.align_ranks(Xc, Yc, xnc, ync, n, 3, True, want_top3=True)
    assert ops.LAST_RANK_INFO["mode"] == "chain"
    np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), np.arange(n))
    np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), np.arange(n))
    np.testing.assert_array_equal(res.top3_idx.cpu().numpy(), np.tile(np.arange(3), (n, 1)))


@pytest.mark.parametrize("n,d,k", [(1, 64, 1), (2, 8, 2), (127, 40, 5), (129, 64, 10), (257, 300, 16), (513, 64, 3)])
def test_ragged_shapes(cuda_device, n, d, k):
    x, y = _clustered(n, d, 1.0, n)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    res = evaluate.align_ranks(X, Y, xn, yn, n, k, True, want_top3=n >= 3)
    ref = oracle.align_eval(x, y, True, k)
    np.testing.assert_array_equal(res.rank_l2r.cpu().numpy(), ref["rank_l2r"])
    np.testing.assert_array_equal(res.rank_r2l.cpu().numpy(), ref["rank_r2l"])


def test_argument_errors(cuda_device):
    x, y = _clustered(64, 64, 1.0, 0)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    with pytest.raises(_lib.SnagError):
        evaluate.align_ranks(X, Y, xn, yn, 64, 17, True)         # more neighbours than the fused path keeps
    with pytest.raises(ValueError):
        evaluate.align_ranks(X[:8], Y[:8], xn[:8], yn[:8], 8, 10, True)   # k > n: torch.topk raises in the reference
    with pytest.raises((ValueError, TypeError)):
        ops.eval_rowtopk(X.float(), Y, xn, yn, 64, 64)           # wrong dtype
    with pytest.raises(ValueError):
        ops.prep_bf16(torch.zeros((0, 64), device=cuda_device), None, True)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_equals_unsharded(cuda_device, world):
    """Targets split over `world` ranks (lockstep-simulated on one GPU, same kernels, same collectives' semantics):
    integer counters and exactly-merged candidate lists make the result bit-identical for any number of ranks."""
    n, d, k = 2300, 320, 10
    x, y = _clustered(n, d, 4.0, 17)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    one = evaluate.align_ranks(X, Y, xn, yn, n, k, True, want_top3=True)
    many = evaluate.simulate_sharded(
        lambda r: evaluate._align_ranks_steps(ops, X, Y, xn, yn, n, k, True, True, world, r), world)
    for res in (many[0], many[-1]):
        assert torch.equal(res.rank_l2r, one.rank_l2r) and torch.equal(res.rank_r2l, one.rank_r2l)
        assert torch.equal(res.nv1, one.nv1) and torch.equal(res.nv2, one.nv2)
        assert torch.equal(res.top3_idx, one.top3_idx)


def test_pairwise_distances_dropin(cuda_device):
    rng = np.random.RandomState(5)
    x = oracle.bf16_round(rng.randn(300, 200).astype(np.float32))
    y = oracle.bf16_round(rng.randn(421, 200).astype(np.float32))
    got = evaluate.pairwise_distances(torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device))
    ref = oracle.pairwise_distances(x, y)
    # rows are not normalised here (||x||^2 ~ 200): the error of d = xn + yn - 2 x.y scales with the magnitude of its
    # terms, not with d (the self-distance diagonal is a cancellation of ~400 - ~400)
    xn, yn = oracle.norm2(x), oracle.norm2(y)
    tol = 3e-6 * (xn[:, None] + yn[None, :])
    assert (np.abs(got.cpu().numpy() - ref) <= tol).all()
    sym = evaluate.pairwise_distances(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    assert (np.abs(sym - oracle.pairwise_distances(x)) <= 3e-6 * (xn[:, None] + xn[None, :])).all()
    assert (sym >= 0).all()                                                       # the clamp of src/utils.py:218


def test_evaluate_alignment_end_to_end(cuda_device):
    """final_emb + index tensors in, metrics out (the Runner._test data flow): gather + normalise + round on device."""
    rng = np.random.RandomState(9)
    N, n, d = 3000, 1200, 300
    emb = rng.randn(N, d).astype(np.float32)
    left = rng.permutation(N // 2)[:n]
    right = N // 2 + rng.permutation(N // 2)[:n]
    emb[right] = emb[left] + 0.8 * rng.randn(n, d).astype(np.float32)
    out = evaluate.evaluate_alignment(torch.from_numpy(emb).to(cuda_device), torch.from_numpy(left).to(cuda_device),
                                      torch.from_numpy(right).to(cuda_device), csls=True, csls_k=10)
    xr = oracle.bf16_round(oracle.normalize_rows(emb[left]))
    yr = oracle.bf16_round(oracle.normalize_rows(emb[right]))
    ref = oracle.align_eval(xr, yr, True, 10)
    # the device normalisation may differ from numpy's by one bf16 ulp on a few elements: compare metrics loosely,
    # ranks on at least 99.5 % of the pairs
    assert (out["ranks"].rank_l2r.cpu().numpy() == ref["rank_l2r"]).mean() > 0.995
    assert abs(out["l2r"].mrr - oracle.metrics(ref["rank_l2r"])["mrr"]) < 2e-3


# ------------------------------------------------------------------------------------------------ BASELINE sizes
@pytest.fixture(scope="module")
def c2_problem(cuda_device):
    """configs[1]-shaped evaluation: 10 500 test pairs, joint width 1800 (fr_en + surface), CSLS k=10."""
    n, d = 10500, 1800
    g = torch.Generator(device="cuda").manual_seed(3408)
    centres = torch.randn((64, d), generator=g, device="cuda")
    x = torch.randn((n, d), generator=g, device="cuda") + centres[torch.randint(0, 64, (n,), generator=g, device="cuda")]
    y = x + 8.0 * torch.randn((n, d), generator=g, device="cuda")
    X, xn = ops.prep_bf16(x, None, True)
    Y, yn = ops.prep_bf16(y, None, True)
    return X, Y, xn, yn, n


def test_full_size_properties(cuda_device, c2_problem):
    X, Y, xn, yn, n = c2_problem
    base = evaluate.align_ranks(X, Y, xn, yn, n, 10, True)
    m = evaluate.metrics_from_ranks(base.rank_l2r)
    assert 0.5 < m.acc[0] <= 1.0 and m.acc[0] <= m.acc[1] <= m.acc[2]
    # determinism
    again = evaluate.align_ranks(X, Y, xn, yn, n, 10, True)
    assert torch.equal(again.rank_l2r, base.rank_l2r) and torch.equal(again.rank_r2l, base.rank_r2l)
    # a common permutation of the pairs permutes the ranks (ties aside: none expected on continuous data)
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    p = evaluate.align_ranks(X[perm].contiguous(), Y[perm].contiguous(), xn[perm].contiguous(), yn[perm].contiguous(),
                             n, 10, True)
    assert (p.rank_l2r == base.rank_l2r[perm]).float().mean() > 0.9995
    assert (p.rank_r2l == base.rank_r2l[perm]).float().mean() > 0.9995
    # swapping the two sides swaps the directions
    s = evaluate.align_ranks(Y, X, yn, xn, n, 10, True)
    assert (s.rank_l2r == base.rank_r2l).float().mean() > 0.9995 and (s.rank_r2l == base.rank_l2r).float().mean() > 0.9995
    # sharded over 4 ranks: identical
    many = evaluate.simulate_sharded(lambda r: evaluate._align_ranks_steps(ops, X, Y, xn, yn, n, 10, True, False, 4, r), 4)
    assert torch.equal(many[2].rank_l2r, base.rank_l2r) and torch.equal(many[2].rank_r2l, base.rank_r2l)


def test_full_size_identity(cuda_device, c2_problem):
    X, _, xn, _, n = c2_problem
    res = evaluate.align_ranks(X, X, xn, xn, n, 10, True)
    assert int(res.rank_l2r.max()) == 0 and int(res.rank_r2l.max()) == 0


def test_full_size_rank_counts_against_materialised(cuda_device, c2_problem):
    """The fused counters against a brute-force count on the materialised fp32 matrix (same kernels' distances)."""
    X, Y, xn, yn, n = c2_problem
    res = evaluate.align_ranks(X, Y, xn, yn, n, 10, True)
    d = ops.sim_write(X, Y, xn, yn, n, n, 1)
    dist = 1 - ((2 * (1 - d) - res.nv1[:, None]) - res.nv2[None, :])
    g = res.g
    lt = (dist < g[:, None]).sum(1).int()
    ltc = (dist < g[None, :]).sum(0).int()
    assert (lt == res.rank_l2r).float().mean() > 0.9995 and (ltc == res.rank_r2l).float().mean() > 0.9995


@pytest.mark.parametrize("n1,n2,k", [(300, 421, 10), (1000, 1000, 3), (257, 96, 16), (2049, 1500, 1)])
def test_csls_sim_dropin_bitwise(cuda_device, n1, n2, k):
    """csls_sim on a materialised matrix (src/utils.py:417-435): selection + largest-first fp32 mean + two fp32 ops,
    so the result is bit-identical to the oracle for the same input matrix."""
    rng = np.random.RandomState(n1 + n2)
    sim = rng.randn(n1, n2).astype(np.float32)
    sim[5, :7] = sim[5, 0]                      # ties inside a row do not matter for a top-k of VALUES
    got = evaluate.csls_sim(torch.from_numpy(sim).to(cuda_device), k)
    ref, nv1, nv2 = oracle.csls_sim(sim, k, return_nv=True)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    _, g1, g2 = ops.csls_sim_matrix(torch.from_numpy(sim).to(cuda_device), k, want_out=False)
    np.testing.assert_array_equal(g1.cpu().numpy(), nv1)
    np.testing.assert_array_equal(g2.cpu().numpy(), nv2)
    with pytest.raises(RuntimeError):
        evaluate.csls_sim(torch.from_numpy(sim[:4, :4].copy()).to(cuda_device), 5)     # k > n, as torch.topk raises


@pytest.mark.parametrize("n1,n2,k", [(300, 421, 17), (700, 650, 64), (1500, 1100, 1000), (64, 2100, 33)])
def test_csls_sim_dropin_large_k_bitwise(cuda_device, n1, n2, k):
    """k beyond the candidate-list length (the reference's torch.topk takes any k): radix selection + sorted
    largest-first sum, bit-identical to the oracle — including rows with many copies of the k-th value."""
    rng = np.random.RandomState(n1 + n2 + k)
    sim = rng.randn(n1, n2).astype(np.float32)
    sim[3, :] = np.float32(0.25)                 # a constant row: every copy of the k-th value
    sim[:, 5] = np.round(sim[:, 5] * 4) / 4      # a column with heavy ties
    sim[7, :40] = -0.0
    got = evaluate.csls_sim(torch.from_numpy(sim).to(cuda_device), k)
    ref, nv1, nv2 = oracle.csls_sim(sim, k, return_nv=True)
    _, g1, g2 = ops.csls_sim_matrix(torch.from_numpy(sim).to(cuda_device), k, want_out=False)
    np.testing.assert_array_equal(g1.cpu().numpy(), nv1)
    np.testing.assert_array_equal(g2.cpu().numpy(), nv2)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)


def test_evaluation_with_large_k_is_the_reference_composition(cuda_device):
    """csls_k > KT no longer raises: evaluate_alignment takes the materialised route and returns what the reference's
    call sequence (main.py:386-429) returns on the same distance matrix."""
    rng = np.random.RandomState(5)
    n, d, k = 900, 200, 40
    emb = torch.from_numpy(rng.randn(2 * n, d).astype(np.float32)).to(cuda_device)
    emb[n:] = emb[:n] + 2.0 * torch.from_numpy(rng.randn(n, d).astype(np.float32)).to(cuda_device)
    left, right = torch.arange(n, device=cuda_device), torch.arange(n, 2 * n, device=cuda_device)
    out = evaluate.evaluate_alignment(emb, left, right, csls=True, csls_k=k)
    assert out["ranks"].info["materialised"]
    fe = torch.nn.functional.normalize(emb)
    distance = evaluate.pairwise_distances(fe[left], fe[right])
    sim = (1 - distance).cpu().numpy()
    ref = 1 - torch.from_numpy(oracle.csls_sim(sim, k))
    want = np.asarray([(torch.sort(ref[i], stable=True)[1] == i).nonzero().item() for i in range(n)], np.int32)
    np.testing.assert_array_equal(out["ranks"].rank_l2r.cpu().numpy(), want)
    want_c = np.asarray([(torch.sort(ref[:, j], stable=True)[1] == j).nonzero().item() for j in range(n)], np.int32)
    np.testing.assert_array_equal(out["ranks"].rank_r2l.cpu().numpy(), want_c)


def test_reference_call_sequence_materialised(cuda_device):
    """main.py:386-393 exactly as the reference writes it, through the two materialising drop-ins."""
    fx = load_golden("eval_n384_d96_k10")
    x = torch.from_numpy(fx["x"]).to(cuda_device)
    y = torch.from_numpy(fx["y"]).to(cuda_device)
    distance = evaluate.pairwise_distances(x, y)
    distance = 1 - evaluate.csls_sim(1 - distance, 10)
    d = distance.cpu()
    ranks = [(torch.sort(d[i], stable=True)[1] == i).nonzero().item() for i in range(d.shape[0])]
    np.testing.assert_array_equal(np.asarray(ranks, np.int32), fx["rank_l2r"])


def test_two_sweep_equals_three_sweep_bitwise(cuda_device):
    """The two-sweep CSLS path (column neighbourhoods collected during the row sweep through sample-derived admission
    thresholds) must reproduce the three-sweep path bit for bit: the candidates are a superset of every column's
    true neighbourhood and the k largest are then selected exactly."""
    n, d, k = evaluate.TWO_SWEEP_MIN_N + 1234, 320, 10
    g = torch.Generator(device="cuda").manual_seed(11)
    centres = torch.randn((64, d), generator=g, device="cuda")
    x = torch.randn((n, d), generator=g, device="cuda") + centres[torch.randint(0, 64, (n,), generator=g, device="cuda")]
    y = x + 4.0 * torch.randn((n, d), generator=g, device="cuda")
    X, xn = ops.prep_bf16(x, None, True)
    Y, yn = ops.prep_bf16(y, None, True)
    del x, y
    a = evaluate.align_ranks(X, Y, xn, yn, n, k, True, two_sweep=False)
    b = evaluate.align_ranks(X, Y, xn, yn, n, k, True, two_sweep=True)
    assert b.launches > a.launches                       # the fused path really ran
    assert torch.equal(a.nv1, b.nv1) and torch.equal(a.nv2, b.nv2)
    assert torch.equal(a.rank_l2r, b.rank_l2r) and torch.equal(a.rank_r2l, b.rank_r2l)
    # sharded (3 ranks, simulated): same result again
    many = evaluate.simulate_sharded(
        lambda r: evaluate._align_ranks_steps(ops, X, Y, xn, yn, n, k, True, False, 3, r, True), 3)
    assert torch.equal(many[1].nv2, a.nv2) and torch.equal(many[1].rank_r2l, a.rank_r2l)
    # candidate statistics: the sample bound leaves ~k*n/m candidates per target, whatever the data
    m, cap = evaluate.two_sweep_plan(n, k)
    sel = torch.randperm(n, generator=torch.Generator(device="cpu").manual_seed(3408))[:m].sort()[0].cuda()
    part_s = ops.eval_rowtopk(Y, X.index_select(0, sel), yn, xn.index_select(0, sel), n, m)
    _, cand_s = ops.topk_merge_mean(part_s, k, want_nv=False, want_cand=True)
    colthr, colb = ops.col_threshold(cand_s, k, yn)
    _, _, stream, srow, scnt = ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, cap)
    cval, cidx, overflow = ops.col_cand_reduce(stream, srow, scnt, n, k)
    assert int(overflow.item()) == 0 and bool((cidx[:, -k:] >= 0).all())
    per_col = float(scnt.sum().item()) / n
    assert abs(per_col / (k * n / m) - 1.0) < 0.1
    nv2 = ops.topk_rescore(Y, X, yn, xn, cidx, cval, k, n, "cols")
    assert torch.equal(nv2, a.nv2) and ops.LAST_TOPK_INFO["cols"]["flagged"] == 0
    # a stream that is too small must be reported, not silently truncated
    _, _, stream, srow, scnt = ops.eval_rowcoltopk(X, Y, xn, yn, n, n, colthr, colb, 1024)
    _, _, overflow = ops.col_cand_reduce(stream, srow, scnt, n, k)
    assert int(overflow.item()) == 1


@pytest.mark.parametrize("n,d,k,csls", [(3000, 1200, 10, True), (1200, 300, 3, True), (900, 96, 5, False)])
def test_lazy_single_sync_path_equals_synchronising_path(cuda_device, n, d, k, csls):
    """The sync-free evaluation (tolerances from the assumed unit norm, counters checked once at the end) returns what
    the synchronising path returns, bit for bit, and reports the same diagnostics."""
    x, y = _clustered(n, d, 6.0, 21)
    X, Y, xn, yn = _prep(x, y, cuda_device)
    a = evaluate.align_ranks(X, Y, xn, yn, n, k, csls, want_top3=True, lazy=False)
    b = evaluate.align_ranks(X, Y, xn, yn, n, k, csls, want_top3=True, lazy=True)
    assert torch.equal(a.rank_l2r, b.rank_l2r) and torch.equal(a.rank_r2l, b.rank_r2l) and torch.equal(a.g, b.g)
    assert torch.equal(a.top3_idx, b.top3_idx)
    if csls:
        assert torch.equal(a.nv1, b.nv1) and torch.equal(a.nv2, b.nv2)
        assert b.info["neighbourhoods"]["rows"]["unverified"] == 0 and b.info["neighbourhoods"]["cols"]["unverified"] == 0
    assert b.info["rank_sweep"]["mode"] == "band" and b.info["rank_sweep"]["deferred"] <= b.info["rank_sweep"]["cap"]
    ref = oracle.align_eval(x, y, csls, k)
    np.testing.assert_array_equal(b.rank_l2r.cpu().numpy(), ref["rank_l2r"])
    np.testing.assert_array_equal(b.rank_r2l.cpu().numpy(), ref["rank_r2l"])
    # rows that are not unit norm break the lazy path's assumption: it must notice and hand over
    X2, xn2 = ops.prep_bf16(torch.from_numpy(2.0 * x).to(cuda_device), None, normalize=False)
    c = evaluate.align_ranks(X2, Y, xn2, yn, n, k, csls)
    ref2 = oracle.align_eval(2.0 * x, y, csls, k)
    np.testing.assert_array_equal(c.rank_l2r.cpu().numpy(), ref2["rank_l2r"])


def test_graphed_evaluation_replays(cuda_device):
    """evaluate_alignment replays one CUDA graph per problem shape: same metrics and ranks as the eager evaluation, also
    after the embeddings and the test pairs change in place between replays."""
    rng = np.random.RandomState(4)
    N, n, d = 5000, 2100, 300
    outs = []
    for trial in range(3):
        emb = rng.randn(N, d).astype(np.float32)
        left = rng.permutation(N // 2)[:n]
        right = N // 2 + rng.permutation(N // 2)[:n]
        emb[right] = emb[left] + 5.0 * rng.randn(n, d).astype(np.float32)   # noisy enough that the three MRRs differ
        args = (torch.from_numpy(emb).to(cuda_device), torch.from_numpy(left).to(cuda_device), torch.from_numpy(right).to(cuda_device))
        g = evaluate.evaluate_alignment(*args, csls=True, csls_k=10, want_top3=True, graph=True)
        e = evaluate.evaluate_alignment(*args, csls=True, csls_k=10, want_top3=True, graph=False)
        assert g["ranks"].info.get("cuda_graph") is True and not e["ranks"].info.get("cuda_graph")
        assert torch.equal(g["ranks"].rank_l2r, e["ranks"].rank_l2r) and torch.equal(g["ranks"].rank_r2l, e["ranks"].rank_r2l)
        assert torch.equal(g["ranks"].top3_idx, e["ranks"].top3_idx) and torch.equal(g["ranks"].nv1, e["ranks"].nv1)
        assert g["l2r"].mrr == e["l2r"].mrr and np.array_equal(g["r2l"].acc, e["r2l"].acc)
        outs.append(g["l2r"].mrr)
    assert len(set(outs)) == 3                      # three different problems went through the same graph
    assert len([k for k in evaluate._GRAPHS if k[1] == n]) == 1


@pytest.mark.parametrize("name", golden_names("l1_"))
def test_l1_distance_golden(cuda_device, name):
    """--distance 1 (main.py:387-390) on the device: distances bit-identical to scipy's (as stored by the reference),
    ranks and prediction ids equal to the reference's loops."""
    fx = load_golden(name)
    x, y = torch.from_numpy(fx["x"]).to(cuda_device), torch.from_numpy(fx["y"]).to(cuda_device)
    np.testing.assert_array_equal(ops.l1_distance(x, y).cpu().numpy(), fx["distance"])
    n = x.shape[0]
    emb = torch.cat([x, y], 0)
    left = torch.arange(0, n, device=cuda_device)
    right = torch.arange(n, 2 * n, device=cuda_device)
    out = evaluate.evaluate_alignment_l1(emb, left, right, csls=bool(fx["csls"]), csls_k=int(fx["k"]), want_top3=True)
    np.testing.assert_array_equal(out["ranks"].rank_l2r.cpu().numpy(), fx["rank_l2r"])
    np.testing.assert_array_equal(out["ranks"].rank_r2l.cpu().numpy(), fx["rank_r2l"])
    np.testing.assert_array_equal(out["ranks"].top3_idx.cpu().numpy(), fx["top3"])


def test_l1_distance_against_oracle_at_c1_width(cuda_device):
    rng = np.random.RandomState(12)
    n, d = 1500, 1200
    x = oracle.normalize_rows(rng.randn(n, d).astype(np.float32))
    y = oracle.normalize_rows((x + 0.05 * rng.randn(n, d)).astype(np.float32))
    emb = torch.from_numpy(np.concatenate([x, y], 0)).to(cuda_device)
    out = evaluate.evaluate_alignment_l1(emb, torch.arange(0, n, device=cuda_device), torch.arange(n, 2 * n, device=cuda_device),
                                         csls=True, csls_k=10)
    xr = torch.nn.functional.normalize(torch.from_numpy(x)).numpy()      # the path re-normalises, as main.py:379 does
    yr = torch.nn.functional.normalize(torch.from_numpy(y)).numpy()
    ref = oracle.align_eval_l1(xr, yr, True, 10)
    agree = (out["ranks"].rank_l2r.cpu().numpy() == ref["rank_l2r"]).mean()
    assert agree > 0.995                 # F.normalize on the device vs the host may differ in the last bit of a few rows
    got = ops.l1_distance(torch.from_numpy(xr).to(cuda_device), torch.from_numpy(yr).to(cuda_device)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.l1_distance(xr, yr))
    r, c = ops.matrix_rank(torch.from_numpy(ref["dist"]).to(cuda_device))
    np.testing.assert_array_equal(r.cpu().numpy(), ref["rank_l2r"])
    np.testing.assert_array_equal(c.cpu().numpy(), ref["rank_r2l"])
