"""CPU: the oracle (oracle/) against the golden vectors produced by the reference's own functions
(tests/golden/gen_golden.py), plus the domain's invariants. This is what pins the checker."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.conftest import golden_names, load_golden


# ------------------------------------------------------------------------------------------------ evaluation
@pytest.mark.parametrize("name", golden_names("eval_"))
def test_eval_matches_reference(name):
    fx = load_golden(name)
    x, y, k, csls = fx["x"], fx["y"], int(fx["k"]), bool(fx["csls"])
    out = oracle.align_eval(x, y, csls=csls, k=k)
    # integer work: bit-exact
    np.testing.assert_array_equal(out["rank_l2r"], fx["rank_l2r"])
    np.testing.assert_array_equal(out["rank_r2l"], fx["rank_r2l"])
    np.testing.assert_array_equal(out["top3"], fx["top3"])
    # floating point: the reference's torch.mm accumulates in fp32 in library order, the oracle in fp64
    np.testing.assert_allclose(out["g"], fx["g"], rtol=0, atol=2e-6)
    if csls:
        np.testing.assert_allclose(out["nv1"], fx["nv1"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out["nv2"], fx["nv2"], rtol=0, atol=1e-6)
    for side in ("l2r", "r2l"):
        m = oracle.metrics(out[f"rank_{side}"])
        np.testing.assert_array_equal(m["acc"], fx[f"acc_{side}"])
        assert m["mr"] == float(fx[f"mr_{side}"])
        assert m["mrr"] == float(fx[f"mrr_{side}"])


@pytest.mark.parametrize("name", ["eval_n384_d96_k10", "eval_ties_dyadic_k4"])
def test_pairwise_and_csls_match_reference_stats(name):
    fx = load_golden(name)
    d = oracle.pairwise_distances(fx["x"], fx["y"])
    np.testing.assert_allclose(np.diag(d), fx["d_diag"], rtol=0, atol=1e-6)
    assert abs(float(d.astype(np.float64).sum()) - float(fx["d_checksum"])) <= 1e-6 * d.size
    sim = (np.float32(1) - d).astype(np.float32)
    _, nv1, nv2 = oracle.csls_sim(sim, int(fx["k"]), return_nv=True)
    np.testing.assert_allclose(nv1, fx["nv1"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(nv2, fx["nv2"], rtol=0, atol=1e-6)


def test_dyadic_fixture_is_exact():
    """With dyadic-rational inputs every op of the chain is exact, so the oracle must match the reference bit for bit."""
    fx = load_golden("eval_ties_dyadic_k4")
    out = oracle.align_eval(fx["x"], fx["y"], csls=True, k=4)
    np.testing.assert_array_equal(out["g"], fx["g"])
    np.testing.assert_array_equal(out["nv1"], fx["nv1"])
    np.testing.assert_array_equal(out["nv2"], fx["nv2"])
    assert float(fx["margin"]) == 0.0          # the fixture really contains exact ties


def test_eval_permutation_invariance():
    fx = load_golden("eval_n384_d96_k10")
    x, y = fx["x"], fx["y"]
    perm = np.random.RandomState(0).permutation(x.shape[0])
    a = oracle.align_eval(x, y, True, 10)
    b = oracle.align_eval(x[perm], y[perm], True, 10)
    np.testing.assert_array_equal(a["rank_l2r"][perm], b["rank_l2r"])
    np.testing.assert_array_equal(a["rank_r2l"][perm], b["rank_r2l"])


def test_eval_identity_is_rank_zero():
    rng = np.random.RandomState(1)
    x = oracle.bf16_round(oracle.normalize_rows(rng.randn(200, 64).astype(np.float32)))
    out = oracle.align_eval(x, x.copy(), True, 5)
    assert out["rank_l2r"].max() == 0 and out["rank_r2l"].max() == 0


def test_k_larger_than_n_raises():
    x = np.eye(4, 8, dtype=np.float32)
    with pytest.raises(RuntimeError):
        oracle.align_eval(x, x, True, 5)
    with pytest.raises(RuntimeError):
        oracle.csls_sim(np.zeros((4, 4), np.float32), 5)


def test_bf16_round_matches_torch():
    rng = np.random.RandomState(2)
    v = np.concatenate([rng.randn(10000).astype(np.float32) * 10.0 ** rng.randint(-6, 6, 10000),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.3895314e38], np.float32)])
    ref = torch.from_numpy(v).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(oracle.bf16_round(v), ref)


def test_dot_is_fp64_accumulated():
    rng = np.random.RandomState(3)
    x = rng.randn(17, 333).astype(np.float32)
    y = rng.randn(29, 333).astype(np.float32)
    ref = (x.astype(np.float64) @ y.astype(np.float64).T).astype(np.float32)
    got = oracle.dot_matrix(x, y)
    # fp64 accumulation in a different order can differ from numpy's by at most one fp32 ulp after rounding
    assert np.max(np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))) <= 1


def test_blocked_dot_is_index_order_fp64():
    """The register-blocked dot kernel of the oracle is a speed-up only: every element equals the scalar fp64
    index-order accumulation rounded once (checked against a plain Python-float loop, which is exactly that)."""
    rng = np.random.RandomState(8)
    x = oracle.bf16_round(rng.randn(9, 77).astype(np.float32))
    y = oracle.bf16_round(rng.randn(31, 77).astype(np.float32))
    got = oracle.dot_matrix(x, y)
    for i in range(9):
        for j in range(31):
            acc = 0.0
            for k in range(77):
                acc += float(x[i, k]) * float(y[j, k])
            assert got[i, j] == np.float32(acc)


@pytest.mark.parametrize("n,d,k,csls", [(701, 96, 10, True), (1030, 64, 3, True), (515, 40, 5, False), (9, 8, 2, True)])
def test_streaming_and_audit_equal_materialised(n, d, k, csls):
    """The streaming evaluation (row blocks, two passes) and the sampled audit restate the same arithmetic as
    align_eval: ranks, neighbourhood means and ground-truth distances are bit-identical, exact ties included."""
    rng = np.random.RandomState(n)
    c = rng.randn(8, d).astype(np.float32)
    x = rng.randn(n, d).astype(np.float32) + c[rng.randint(0, 8, n)]
    y = x + 1.5 * rng.randn(n, d).astype(np.float32)
    x[5], y[7] = x[4], y[6]                               # duplicated rows: exact ties
    x, y = oracle.bf16_round(oracle.normalize_rows(x)), oracle.bf16_round(oracle.normalize_rows(y))
    a = oracle.align_eval(x, y, csls, k)
    for block_rows in (4, 100, 5000):
        b = oracle.align_eval_stream(x, y, csls, k, block_rows)
        for key in b:
            np.testing.assert_array_equal(a[key], b[key], err_msg=f"{key} block_rows={block_rows}")
    sel = rng.permutation(n)[:min(n, 37)]
    r = oracle.audit(x[sel], y, sel, csls, k, a.get("nv2"), False)
    np.testing.assert_array_equal(r["rank"], a["rank_l2r"][sel])
    np.testing.assert_array_equal(r["g"], a["g"][sel])
    c_ = oracle.audit(y[sel], x, sel, csls, k, a.get("nv1"), True)
    np.testing.assert_array_equal(c_["rank"], a["rank_r2l"][sel])
    np.testing.assert_array_equal(c_["g"], a["g"][sel])
    if csls:
        np.testing.assert_array_equal(r["nv"], a["nv1"][sel])
        np.testing.assert_array_equal(c_["nv"], a["nv2"][sel])


# ------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("name", golden_names("icl_"))
def test_icl_matches_reference(name):
    fx = load_golden(name)
    w = fx["weight_norm"] if int(fx["weighted"]) else None
    loss = oracle.icl_loss(fx["emb"], fx["links"], float(fx["tau"]), float(fx["ab_weight"]), w)
    np.testing.assert_allclose(loss, fx["loss"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("name", golden_names("ial_"))
def test_ial_matches_reference(name):
    fx = load_golden(name)
    loss = oracle.ial_loss(fx["src"], fx["tar"], fx["links"], float(fx["tau"]), float(fx["ab_weight"]), float(fx["zoom"]),
                           str(fx["reduction"]))
    np.testing.assert_allclose(loss, fx["loss"], rtol=2e-4, atol=1e-9)


def test_ial_identity_is_zero():
    fx = load_golden("ial_tau0.5_mean")
    assert abs(float(oracle.ial_loss(fx["src"], fx["src"], fx["links"], 0.5))) < 1e-9


def test_icl_equals_cross_entropy_form():
    """ICL == alpha*CE([ab|aa\\diag]) + (1-alpha)*CE([ba|bb\\diag]) with arange labels."""
    fx = load_golden("icl_tau0.1_nw")
    emb, links = torch.from_numpy(fx["emb"]), fx["links"].astype(np.int64)
    z = torch.nn.functional.normalize(emb)
    a, b = z[links[:, 0]], z[links[:, 1]]
    bsz = a.shape[0]
    eye = torch.eye(bsz) * 1e9
    la = torch.cat([a @ b.t(), a @ a.t() - eye], 1) / 0.1
    lb = torch.cat([b @ a.t(), b @ b.t() - eye], 1) / 0.1
    lab = torch.arange(bsz)
    ce = 0.5 * torch.nn.functional.cross_entropy(la, lab) + 0.5 * torch.nn.functional.cross_entropy(lb, lab)
    np.testing.assert_allclose(oracle.icl_loss(fx["emb"], fx["links"], 0.1), ce.item(), rtol=1e-5)


def test_multi_loss_layer_matches_reference():
    fx = load_golden("mll")
    np.testing.assert_allclose(oracle.multi_loss_layer(fx["losses"], fx["log_vars"]), fx["out"], rtol=1e-6)


# ------------------------------------------------------------------------------------------------ noise
@pytest.mark.parametrize("name", golden_names("noise_"))
def test_noise_matches_reference_bitwise(name):
    fx = load_golden(name)
    out = oracle.add_noise_to_embeddings(fx["x"], fx["mean"], fx["std"], fx["mask"], fx["z"], float(fx["mask_ratio"]))
    np.testing.assert_array_equal(out, fx["out"])
    np.testing.assert_array_equal(out[~fx["mask"]], fx["x"][~fx["mask"]])        # untouched rows are bit-identical


def test_rowblend_matches_reference_bitwise():
    fx = load_golden("rowblend")
    np.testing.assert_array_equal(oracle.rowblend(fx["e"], fx["noise"], fx["mask"], float(fx["mask_ratio"])), fx["out"])


def test_noise_identities():
    rng = np.random.RandomState(4)
    x = rng.randn(50, 8).astype(np.float32)
    mean, std = oracle.col_mean_std(x)
    none = np.zeros(50, bool)
    allm = np.ones(50, bool)
    z = rng.randn(50, 8).astype(np.float32)
    np.testing.assert_array_equal(oracle.add_noise_to_embeddings(x, mean, std, none, z[:0], 0.7), x)      # r = 0
    np.testing.assert_array_equal(oracle.add_noise_to_embeddings(x, mean, std, allm, z, 0.0), x)          # rho = 0
    out = oracle.add_noise_to_embeddings(x, mean, np.zeros_like(std), allm, z, 1.0)                        # rho = 1, std = 0
    np.testing.assert_array_equal(out, np.broadcast_to(mean, x.shape))
    tm, ts = torch.from_numpy(x).mean(0).numpy(), torch.from_numpy(x).std(0).numpy()
    np.testing.assert_allclose(mean, tm, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(std, ts, rtol=1e-6, atol=1e-7)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    r = oracle.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(v) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = oracle.philox4x32_10(np.uint32(0xFFFFFFFF), np.uint32(0xFFFFFFFF), np.uint32(0xFFFFFFFF), np.uint32(0xFFFFFFFF),
                             0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = oracle.philox4x32_10(np.uint32(0x243F6A88), np.uint32(0x85A308D3), np.uint32(0x13198A2E), np.uint32(0x03707344),
                             0xA4093822, 0x299F31D0)
    assert [int(v) for v in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


# ------------------------------------------------------------------------------------------------ link mining
@pytest.mark.parametrize("name", golden_names("mining_"))
def test_mutual_nn_matches_reference(name):
    """oracle.mutual_nn / iter_new_links against what SNAG.Iter_new_links returned (model/SNAG.py:192-208)."""
    fx = load_golden(name)
    emb, left, right = fx["emb"], fx["left"].tolist(), fx["right"].tolist()
    pl, pr, dl, dr = oracle.mutual_nn(emb[left], emb[right])
    np.testing.assert_array_equal(pl, fx["preds_l"])
    np.testing.assert_array_equal(pr, fx["preds_r"])
    np.testing.assert_allclose(dl, fx["dmin_l"], atol=1e-6, rtol=0)
    got = oracle.iter_new_links(left, right, emb, [], True)
    assert got == [tuple(t) for t in fx["links_refresh"].tolist()]
    got = oracle.iter_new_links(left, right, emb, [tuple(t) for t in fx["prev"].tolist()], False)
    assert got == [tuple(t) for t in fx["links_filter"].tolist()]


@pytest.mark.parametrize("name", golden_names("mining_"))
def test_mining_host_logic_on_cpu_backend(name):
    """snag_b200.mining (sample pre-pass bound, list merge, packed column keys, link filter) with the oracle standing in
    for the kernels."""
    import torch
    from snag_b200 import mining
    from tests import oracle_backend
    fx = load_golden(name)
    emb = torch.from_numpy(fx["emb"])
    left, right = fx["left"].tolist(), fx["right"].tolist()
    pl, pr, dl, dr = mining.mutual_nearest(emb[left], emb[right], _backend=oracle_backend)
    np.testing.assert_array_equal(pl.numpy(), fx["preds_l"])
    np.testing.assert_array_equal(pr.numpy(), fx["preds_r"])
    np.testing.assert_allclose(dr.numpy(), fx["dmin_r"], atol=1e-6, rtol=0)
    assert mining.iter_new_links(left, right, emb, [], True, _backend=oracle_backend) == [tuple(t) for t in fx["links_refresh"].tolist()]
    prev = [tuple(t) for t in fx["prev"].tolist()]
    assert mining.iter_new_links(left, right, emb, prev, False, _backend=oracle_backend) == [tuple(t) for t in fx["links_filter"].tolist()]


@pytest.mark.parametrize("name", golden_names("fusion_"))
def test_joint_fuse_matches_reference(name):
    """model/SNAG_tools.py:44-49 as run by the reference's own MformerFusion (gen_golden_fusion.py)."""
    fx = load_golden(name)
    M = int(fx["M"])
    j, fz = oracle.joint_fuse([fx[f"emb{m}"] for m in range(M)], fx["weight_norm"], fx["weight_norm_fz"])
    np.testing.assert_allclose(j, fx["joint"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(fz, fx["joint_fz"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("seeds_"))
def test_seed_induction_matches_reference(name):
    """src/data.py:367-402 as run by the reference itself (gen_golden_seeds.py)."""
    fx = load_golden(name)
    links = oracle.visual_pivot_induction(fx["left"].tolist(), fx["right"].tolist(), fx["feats"], int(fx["unsup_k"]))
    np.testing.assert_array_equal(links, fx["links"])


@pytest.mark.parametrize("name", golden_names("l1_"))
def test_l1_distance_path_matches_reference(name):
    """--distance 1: the oracle's cityblock distances equal scipy's cdist result as the reference stores it (fp32), and
    CSLS + ranks + top-3 on them equal what the reference's own csls_sim and ranking loops produced."""
    fx = load_golden(name)
    got = oracle.l1_distance(fx["x"], fx["y"])
    np.testing.assert_array_equal(got, fx["distance"])
    out = oracle.align_eval_l1(fx["x"], fx["y"], bool(fx["csls"]), int(fx["k"]))
    # L1 distances of unit rows are O(10): a few fp32 ulps (1e-6 each) from torch.mean's summation order in csls_sim
    np.testing.assert_allclose(out["dist"], fx["dist"], rtol=1e-6, atol=2e-6)
    np.testing.assert_array_equal(out["rank_l2r"], fx["rank_l2r"])
    np.testing.assert_array_equal(out["rank_r2l"], fx["rank_r2l"])
    np.testing.assert_array_equal(out["top3"], fx["top3"])
